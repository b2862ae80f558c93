// Does a MUFU burst of one warp block the OTHER warps' MIO instructions (shared-memory loads, hence presumably tcgen05.ld/st,
// mbarrier waits) on the same scheduler?  One CTA, 8 warps: warps 0-3 (one per scheduler) run pattern A, warps 4-7 pattern B.
//   A: 0 idle | 1 bursts: 16 x tanh.approx back to back, then 96 dependent-free HFMA2 | 2 interleaved: (1 tanh + 6 HFMA2) x 16
//   B: 0 LDS loop (8 independent ld.shared.v4 per iteration) | 1 HFMA2 loop (control: FMA pipe only) | 2 tanh loop (8 per iteration)
// Prints cycles per B iteration for every (A, B).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mio_hol mio_hol.cu
#include <cstdio>
#include <cstdint>
__global__ void __launch_bounds__(256, 1) k(int pa, int pb, long long* out, float* sink) {
    __shared__ float4 sm[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sm[threadIdx.x] = make_float4(lane, 1, 2, 3);
    __syncthreads();
    float x[16]; uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 0.01f * (lane + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0x38003800u + i;
    const uint32_t m = 0x3bff3bffu;
    long long t0 = 0, t1 = 0;
    if (warp < 4) {                                   // pattern A: runs for a fixed number of iterations, longer than B
        if (pa == 0) { __syncthreads(); return; }
        for (int it = 0; it < 6000; ++it) {
            if (pa == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
                for (int r = 0; r < 12; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(m));
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
                    for (int j = 0; j < 6; ++j) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[(i + j) & 7]) : "r"(m));
                }
            }
        }
        float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
        uint32_t q = 0; for (int i = 0; i < 8; ++i) q ^= h[i];
        sink[threadIdx.x] = s + q;
        __syncthreads();
        return;
    }
    float4 acc = make_float4(0, 0, 0, 0);
    t0 = clock64();
    for (int it = 0; it < 2000; ++it) {
        if (pb == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"((uint32_t)__cvta_generic_to_shared(&sm[(lane + 32 * i) & 255])));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        } else if (pb == 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(m));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    t1 = clock64();
    float s = acc.x + acc.y + acc.z + acc.w; for (int i = 0; i < 16; ++i) s += x[i];
    uint32_t q = 0; for (int i = 0; i < 8; ++i) q ^= h[i];
    sink[threadIdx.x] = s + q;
    if (lane == 0) out[warp - 4] = (t1 - t0) / 2000;
    __syncthreads();
}
int main() {
    long long* out; float* sink;
    cudaMallocManaged(&out, 4 * sizeof(long long)); cudaMalloc(&sink, 256 * 4);
    const char* an[3] = {"A idle", "A bursts (16 tanh, 96 hfma2)", "A interleaved (tanh + 6 hfma2) x16"};
    const char* bn[3] = {"B: 8 x LDS.128", "B: 32 x HFMA2", "B: 8 x tanh"};
    for (int pb = 0; pb < 3; ++pb)
        for (int pa = 0; pa < 3; ++pa) {
            k<<<1, 256>>>(pa, pb, out, sink);
            cudaDeviceSynchronize();
            printf("%-18s | %-36s : %lld cycles per B iteration (schedulers 0..3: %lld %lld %lld %lld)\n", bn[pb], an[pa],
                   (out[0] + out[1] + out[2] + out[3]) / 4, out[0], out[1], out[2], out[3]);
        }
    return 0;
}
