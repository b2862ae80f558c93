// Where do the ~235 cycles of "MMA issue" and the ~110 cycles of "issue -> wake" go?  (sm_100a, one CTA per SM)
// One warp, one elected lane: clock64 around the first tcgen05.mma, the remaining k-1, the commit; the whole warp
// then waits on the mbarrier.  Also: commit with no MMA in flight, mbarrier arrive -> wake between two warps,
// bar.sync(128) cost, tcgen05.wait::st after 6 stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu && ./mma_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, int acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, int acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int N, bool d_f32) {
    return (d_f32 ? (1u << 4) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Res { long long c[16]; };
constexpr int REPS = 64;

// test 0: k MMAs (.ts) + commit, split timers.  test 1: same with A from shared memory (.ss).
// test 2: mbarrier arrive (warp 1) -> wake (warp 0).  test 3: bar.sync 128.  test 4: 6 x st.x8 + wait::st
// test 5: full worker-side round without math: wake -> fence -> ld x16 x3 -> wait::ld
__global__ void __launch_bounds__(128, 1) probe(int test, int k, int N, Res* out) {
    __shared__ __align__(128) unsigned char wsm[16384];
    __shared__ unsigned long long bars[4];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(wsm)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t tw = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t v[8] = {0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u};
    tmem_st8(tw + 256, v); tmem_st8(tw + 264, v);
    tc_wait_st(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint64_t b = make_b_desc(smem_u32(wsm), 512, 128);
    const uint64_t adesc = make_b_desc(smem_u32(wsm) + 8192, 2048, 128);
    const uint32_t idesc = make_idesc(N, true);
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t par = 0;
    if (test == 0 || test == 1) {
        if (warp == 0) {
            for (int rep = 0; rep < REPS; ++rep) {
                long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                tc_fence_after();
                const long long cs = clock64();
                if (elect_one()) {
                    c0 = clock64();
                    if (k > 0) { if (test == 0) mma_ts(tb, tb + 256, b, idesc, 0); else mma_ss(tb, adesc, b, idesc, 0); }
                    c1 = clock64();
                    for (int i = 1; i < k; ++i) { if (test == 0) mma_ts(tb + 32 * (i & 3), tb + 256 + 8 * (i & 1), b, idesc, i >> 2); else mma_ss(tb + 32 * (i & 3), adesc, b, idesc, i >> 2); }
                    c2 = clock64();
                    tc_commit(smem_u32(&bars[0]));
                    c3 = clock64();
                }
                __syncwarp();
                const long long c4 = clock64();
                mbar_wait(smem_u32(&bars[0]), par); par ^= 1u;
                const long long c5 = clock64();
                c0 = __shfl_sync(0xffffffffu, c0, __ffs(__ballot_sync(0xffffffffu, c0 != 0)) - 1);
                c1 = __shfl_sync(0xffffffffu, c1, __ffs(__ballot_sync(0xffffffffu, c1 != 0)) - 1);
                c2 = __shfl_sync(0xffffffffu, c2, __ffs(__ballot_sync(0xffffffffu, c2 != 0)) - 1);
                c3 = __shfl_sync(0xffffffffu, c3, __ffs(__ballot_sync(0xffffffffu, c3 != 0)) - 1);
                acc[0] += c0 - cs; acc[1] += c1 - c0; acc[2] += c2 - c1; acc[3] += c3 - c2; acc[4] += c4 - c3; acc[5] += c5 - c4;
                acc[6] += c5 - cs;
            }
            if (lane == 0) for (int i = 0; i < 7; ++i) out->c[i] = acc[i] / REPS;
        }
    } else if (test == 2) {
        // warp 1 arrives at a time stamp it publishes; warp 0 measures wake time - arrive time (same SM clock)
        __shared__ long long stamp;
        for (int rep = 0; rep < REPS; ++rep) {
            __syncthreads();
            if (warp == 1) {
                for (int spin = 0; spin < 200; ++spin) asm volatile("nanosleep.u32 20;");
                if (lane == 0) { stamp = clock64(); mbar_arrive(smem_u32(&bars[1])); }
            } else if (warp == 0) {
                mbar_wait(smem_u32(&bars[1]), par);
                const long long c = clock64();
                acc[0] += c - *(volatile long long*)&stamp;
            }
            par ^= 1u;
        }
        if (warp == 0 && lane == 0) out->c[0] = acc[0] / REPS;
    } else if (test == 3) {
        const long long c0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) asm volatile("bar.sync 1, 128;" ::: "memory");
        const long long c1 = clock64();
        if (threadIdx.x == 0) out->c[0] = (c1 - c0) / REPS;
    } else if (test == 4) {
        long long t_st = 0, t_w = 0, t_f = 0;
        for (int rep = 0; rep < REPS; ++rep) {
            const long long c0 = clock64();
            for (int i = 0; i < k; ++i) tmem_st8(tw + 256 + 8 * (i & 7), v);
            const long long c1 = clock64();
            tc_wait_st();
            const long long c2 = clock64();
            tc_fence_before();
            const long long c3 = clock64();
            t_st += c1 - c0; t_w += c2 - c1; t_f += c3 - c2;
        }
        if (threadIdx.x == 0) { out->c[0] = t_st / REPS; out->c[1] = t_w / REPS; out->c[2] = t_f / REPS; }
    } else if (test == 6 || test == 7) {
        // same bytes as 6 x st.x8: 3 x st.x16 (test 6) or 12 x st.x4 (test 7); k warps active
        long long t_st = 0, t_w = 0;
        if (warp < k) {
            for (int rep = 0; rep < REPS; ++rep) {
                const long long c0 = clock64();
                if (test == 6) { tmem_st16(tw + 256, v); tmem_st16(tw + 272, v); tmem_st16(tw + 288, v); }
                else { for (int i = 0; i < 12; ++i) tmem_st4(tw + 256 + 4 * i, v); }
                const long long c1 = clock64();
                tc_wait_st();
                const long long c2 = clock64();
                t_st += c1 - c0; t_w += c2 - c1;
            }
            if (threadIdx.x == 0) { out->c[0] = t_st / REPS; out->c[1] = t_w / REPS; }
        }
    } else if (test == 8) {
        // 6 x st.x8 with k warps active (test 4 has all four)
        long long t_st = 0, t_w = 0;
        if (warp < k) {
            for (int rep = 0; rep < REPS; ++rep) {
                const long long c0 = clock64();
                for (int i = 0; i < 6; ++i) tmem_st8(tw + 256 + 8 * i, v);
                const long long c1 = clock64();
                tc_wait_st();
                const long long c2 = clock64();
                t_st += c1 - c0; t_w += c2 - c1;
            }
            if (threadIdx.x == 0) { out->c[0] = t_st / REPS; out->c[1] = t_w / REPS; }
        }
    } else if (test == 5) {
        long long t_f = 0, t_ld = 0;
        uint32_t a[16], c[16], d[16], s = 0;
        for (int rep = 0; rep < REPS; ++rep) {
            const long long c0 = clock64();
            tc_fence_after();
            const long long c1 = clock64();
            tmem_ld16(tw, a); if (k > 1) tmem_ld16(tw + 32, c); if (k > 2) tmem_ld16(tw + 64, d);
            tc_wait_ld();
            s += a[0] + (k > 1 ? c[0] : 0) + (k > 2 ? d[0] : 0);
            const long long c2 = clock64();
            t_f += c1 - c0; t_ld += c2 - c1;
        }
        if (threadIdx.x == 0) { out->c[0] = t_f / REPS; out->c[1] = t_ld / REPS; out->c[2] = s; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    }
}

int main() {
    Res* d; cudaMalloc(&d, sizeof(Res));
    Res h;
    auto run = [&](int test, int k, int N) {
        cudaMemset(d, 0, sizeof(Res));
        probe<<<1, 128>>>(test, k, N, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test %d k %d: %s\n", test, k, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    };
    for (int test = 0; test < 2; ++test)
        for (int N : {32, 16})
            for (int k : {0, 1, 2, 4, 8}) {
                run(test, k, N);
                printf("%s N=%2d k=%d: elect %lld | mma#1 %lld | rest %lld | commit %lld | syncwarp %lld | wait %lld | total %lld\n",
                       test == 0 ? "mma.ts" : "mma.ss", N, k, h.c[0], h.c[1], h.c[2], h.c[3], h.c[4], h.c[5], h.c[6]);
            }
    run(2, 0, 32); printf("mbarrier arrive -> peer wake: %lld cyc\n", h.c[0]);
    run(3, 0, 32); printf("bar.sync(128): %lld cyc\n", h.c[0]);
    for (int k : {1, 3, 6}) { run(4, k, 32); printf("st.x8 x%d: issue %lld | wait::st %lld | fence::before %lld\n", k, h.c[0], h.c[1], h.c[2]); }
    for (int k : {1, 4}) { run(8, k, 32); printf("6 x st.x8  (%d warps): issue %lld | wait::st %lld\n", k, h.c[0], h.c[1]); }
    for (int k : {1, 4}) { run(6, k, 32); printf("3 x st.x16 (%d warps): issue %lld | wait::st %lld\n", k, h.c[0], h.c[1]); }
    for (int k : {1, 4}) { run(7, k, 32); printf("12 x st.x4 (%d warps): issue %lld | wait::st %lld\n", k, h.c[0], h.c[1]); }
    for (int k : {1, 2, 3}) { run(5, k, 32); printf("fence::after %lld | ld.x16 x%d + wait::ld %lld\n", h.c[0], k, h.c[1]); }
    return 0;
}
