// How fast can tcgen05.mma instructions be ISSUED on one SM?  (sm_100a; one CTA, 4 warps)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
// W issuing warps (1, 2, 4; warp w on scheduler w), each elected lane issues `kPer` back-to-back M128 x N x K16
// kind::f16 MMAs (.ts form, compile-time operand offsets, distinct accumulators per warp) and one commit, REPS times;
// all warps start together.  Reported: cycles per MMA seen by one issuer and aggregate MMAs per 1000 cycles.
// If the aggregate rate does not grow with W the front end is an SM-wide serial resource; if cycles per MMA do not
// depend on N the cost is per instruction, not per MAC.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@!p bra W_%=;\n\t}"
                 ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int ACC>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "n"(ACC) : "memory");
}
template <int ACC>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "n"(ACC) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int N, bool d_f32) {
    return (d_f32 ? (1u << 4) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Res { long long issue[4], total[4]; };
constexpr int REPS = 32;

// 8 MMAs like one hidden round of the sampler kernel: u (2), z (4), v (2); distinct D per role, A operands at +96..+127
template <int N, bool SS>
__device__ __forceinline__ void round8(uint32_t tg, uint64_t b, uint64_t adesc, uint32_t idesc) {
    if (SS) {
        mma_ss<0>(tg + 32, adesc, b, idesc); mma_ss<1>(tg + 32, adesc, b + 64, idesc);
        mma_ss<0>(tg, adesc, b, idesc); mma_ss<1>(tg, adesc, b + 64, idesc);
        mma_ss<1>(tg, adesc, b + 128, idesc); mma_ss<1>(tg, adesc, b + 192, idesc);
        mma_ss<0>(tg + 64, adesc, b, idesc); mma_ss<1>(tg + 64, adesc, b + 64, idesc);
    } else {
        mma_ts<0>(tg + 32, tg + 96, b, idesc); mma_ts<1>(tg + 32, tg + 104, b + 64, idesc);
        mma_ts<0>(tg, tg + 112, b, idesc); mma_ts<1>(tg, tg + 120, b + 64, idesc);
        mma_ts<1>(tg, tg + 112, b + 128, idesc); mma_ts<1>(tg, tg + 120, b + 192, idesc);
        mma_ts<0>(tg + 64, tg + 96, b, idesc); mma_ts<1>(tg + 64, tg + 104, b + 64, idesc);
    }
}

template <int N, bool SS>
__global__ void __launch_bounds__(128, 1) probe(int n_issuers, int rounds_per_commit, Res* out) {
    __shared__ __align__(128) unsigned char wsm[32768];
    __shared__ unsigned long long bars[4];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(wsm)[i] = 0x2c002c00u;
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
    for (int g = 0; g < 4; ++g) for (int c = 96; c < 128; c += 8) tmem_st8(tw + g * 128 + c, v);
    tc_wait_st(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint64_t b = make_b_desc(smem_u32(wsm), N * 16, 128);
    const uint64_t adesc = make_b_desc(smem_u32(wsm) + 16384, 2048, 128);
    const uint32_t idesc = make_idesc(N, true);
    const uint32_t tg = tb + warp * 128;          // this warp's tile (MMA addresses: lane field 0)
    long long t_issue = 0, t_total = 0;
    uint32_t par = 0;
    if (warp < n_issuers) {
        for (int rep = 0; rep < REPS; ++rep) {
            asm volatile("bar.sync 1, %0;" ::"r"(32 * n_issuers) : "memory");       // issuers start together
            long long c1 = 0;
            const long long c0 = clock64();
            if (elect_one()) {
                tc_fence_after();
                for (int r = 0; r < rounds_per_commit; ++r) round8<N, SS>(tg, b, adesc, idesc);
                c1 = clock64();
                tc_commit(smem_u32(&bars[warp]));
            }
            __syncwarp();
            mbar_wait(smem_u32(&bars[warp]), par); par ^= 1u;
            const long long c2 = clock64();
            c1 = __shfl_sync(0xffffffffu, c1, __ffs(__ballot_sync(0xffffffffu, c1 != 0)) - 1);
            t_issue += c1 - c0; t_total += c2 - c0;
        }
        if (lane == 0) { out->issue[warp] = t_issue / REPS; out->total[warp] = t_total / REPS; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    }
}

template <int N, bool SS>
static void run_all(const char* name) {
    Res* d; Res h;
    cudaMalloc(&d, sizeof(Res));
    for (int rounds : {1, 4}) {
        for (int w : {1, 2, 4}) {
            cudaMemset(d, 0, sizeof(Res));
            probe<N, SS><<<1, 128>>>(w, rounds, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
            cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long worst_issue = 0, worst_total = 0;
            for (int i = 0; i < w; ++i) { if (h.issue[i] > worst_issue) worst_issue = h.issue[i]; if (h.total[i] > worst_total) worst_total = h.total[i]; }
            const int mmas = 8 * rounds;
            printf("%s N=%3d  issuers=%d  %2d MMAs each: issue %5lld cyc (%5.1f / MMA)  issue->all done %5lld cyc  aggregate %.2f MMAs / 1000 cyc\n",
                   name, N, w, mmas, worst_issue, (double)worst_issue / mmas, worst_total, 1000.0 * w * mmas / (double)worst_total);
        }
    }
    cudaFree(d);
}

int main() {
    run_all<16, false>("ts");
    run_all<32, false>("ts");
    run_all<64, false>("ts");
    run_all<128, false>("ts");
    run_all<32, true>("ss");
    run_all<128, true>("ss");
    return 0;
}
