// Micro-benchmarks that size the histogram design: shared-memory atomic throughput with the
// lane-owns-a-bank layout, against LDS/STS read-modify-write and plain byte loads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(const uint8_t* __restrict__ in, uint32_t* out, int iters)
{
    extern __shared__ uint32_t sm[];          // 80 KB hist + 16 KB data
    uint32_t* hist = sm;
    uint8_t* data = (uint8_t*)(sm + 5 * 4096);
    for (int i = threadIdx.x; i < 5 * 4096; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) data[i] = in[(blockIdx.x * 16384 + i) & ((1 << 24) - 1)];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        // each warp walks "records" of 160 bytes: 5 chunks x 32 lanes
        #pragma unroll 1
        for (int r = warp; r < 96; r += 16) {
            const uint8_t* p = data + r * 160 + lane;
            uint32_t* hp = hist + lane;
            #pragma unroll
            for (int c = 0; c < 5; ++c) {
                uint32_t b = p[c * 32] & 127;
                if (MODE == 0) atomicAdd(hp + c * 4096 + (b << 5), 1u);                 // conflict-free ATOMS
                else if (MODE == 1) { uint32_t* q = hp + c * 4096 + (b << 5); *q = *q + 1; }  // LDS+STS (racy; timing only)
                else if (MODE == 2) acc += b;                                             // byte loads only
                else if (MODE == 3) atomicAdd(hist + ((c * 32 + lane) * 128 + b) , 1u);   // [pos][byte] layout: random banks
            }
        }
    }
    if (acc == 0xFFFFFFFF) out[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = hist[threadIdx.x] + acc;
}

template <int MODE> float run(const uint8_t* in, uint32_t* out, int grid, int iters, const char* name)
{
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 512, 98304>>>(in, out, 2);
    cudaEventRecord(a);
    k<MODE><<<grid, 512, 98304>>>(in, out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double bytes = (double)grid * iters * 96 * 160;
    printf("%-28s grid %4d  %8.3f ms  %8.1f G byte-updates/s  (%.2f cycles/warp-instr/SM @1.9GHz, 2 CTA/SM)\n", name, grid, ms,
           bytes / ms / 1e6, 1.9e9 * (ms * 1e-3) / ((double)iters * 96 * 5 * (grid / 148.0)));
    return ms;
}

int main()
{
    uint8_t* in; uint32_t* out;
    cudaMalloc(&in, 1 << 24); cudaMalloc(&out, 1 << 16);
    uint8_t* h = (uint8_t*)malloc(1 << 24);
    uint32_t s = 12345;
    for (int i = 0; i < (1 << 24); ++i) { s = s * 1664525u + 1013904223u; h[i] = 35 + ((s >> 16) % 40); }
    cudaMemcpy(in, h, 1 << 24, cudaMemcpyHostToDevice);
    for (int grid : {148, 296}) {
        run<0>(in, out, grid, 2000, "ATOMS lane-owns-bank");
        run<1>(in, out, grid, 2000, "LDS+IADD+STS lane-owns-bank");
        run<2>(in, out, grid, 2000, "LDS.U8 only");
        run<3>(in, out, grid, 2000, "ATOMS [pos][byte] layout");
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
