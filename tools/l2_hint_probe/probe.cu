// L2 residency probe (B200): does a "hot" buffer written with evict_last survive a stream of evict_first
// traffic larger than the 126 MB L2?  Run under
//   ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
// and read the dram bytes of the `read_hot` launches (and of `write_hot` for the overwrite-in-place case).
//   probe <hot_MB> <stream_MB> <hot_policy 0|1> <stream_policy 0|1> [reps]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pol_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long pol_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_hint(const float4* a, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint(float4* a, float4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

__global__ void write_hot(float4* x, size_t n, int hint, float seed) {
    const unsigned long long pol = pol_last();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = make_float4(seed, i * 1e-9f, 2.f, 3.f);
        if (hint) st_hint(x + i, v, pol); else x[i] = v;
    }
}
__global__ void stream_copy(const float4* a, float4* b, size_t n, int hint) {
    const unsigned long long pol = pol_first();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (hint) st_hint(b + i, ld_hint(a + i, pol), pol); else b[i] = a[i];
    }
}
__global__ void read_hot(const float4* x, size_t n, int hint, float* out) {
    const unsigned long long pol = pol_last();
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = hint ? ld_hint(x + i, pol) : x[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *out = acc;
}

int main(int argc, char** argv) {
    const size_t hot_mb = argc > 1 ? atoi(argv[1]) : 32, str_mb = argc > 2 ? atoi(argv[2]) : 512;
    const int hp = argc > 3 ? atoi(argv[3]) : 1, sp = argc > 4 ? atoi(argv[4]) : 1, reps = argc > 5 ? atoi(argv[5]) : 3;
    const size_t nh = hot_mb * (1 << 20) / 16, ns = str_mb * (1 << 20) / 16;
    float4 *x, *a, *b; float* out;
    cudaMalloc(&x, nh * 16); cudaMalloc(&a, ns * 16); cudaMalloc(&b, ns * 16); cudaMalloc(&out, 4);
    cudaMemset(a, 0, ns * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float t_read = 0.f;
    for (int r = 0; r < reps; ++r) {
        write_hot<<<592, 256>>>(x, nh, hp, (float)r);            // (over)write the hot buffer in place
        stream_copy<<<1184, 256>>>(a, b, ns, sp);               // streaming traffic in between
        cudaEventRecord(e0);
        read_hot<<<592, 256>>>(x, nh, hp, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); t_read = ms;
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("hot %zu MB (policy %d) stream 2x%zu MB (policy %d): read_hot %.1f us = %.0f GB/s  [%s]\n", hot_mb, hp, str_mb, sp,
           t_read * 1e3, hot_mb * 1.048576 / t_read, cudaGetErrorString(e));
    return 0;
}
