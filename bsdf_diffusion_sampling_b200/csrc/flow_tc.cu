// PREC_TC16 path: the per-query flow on 5th-generation tensor cores (tcgen05) for sm_100a.
//
// One persistent CTA per SM, kGroups worker groups of 128 threads (4 warps; thread <-> TMEM lane <->
// query row).  Each group owns one 128-query tile at a time and runs the WHOLE T-step flow for it on
// chip; the groups are independent pipelines that overlap each other's tensor-core round trips:
//
//   per tile:  PE5(wi) -> fp16 -> TMEM A1[:,8:30];  base net, x0 (Philox or replay), p0   (fp32)
//   per step:  state (hi/lo fp16 split) -> TMEM A1[:,0:8]
//              bar.sync(group) ; elected thread:  D_z  = A1 . W1^T                 (K=32, N=32)
//              tcgen05.ld D_z ; SiLU / SiLU' ; tangent seeds -> fp16 -> TMEM A_h, A_u, A_v
//              bar.sync(group) ; elected thread:  D_z,D_u,D_v = A_{h,u,v} . Wl^T   (x (n_hidden-1))
//              ...
//              bar.sync(group) ; elected thread:  D = A . Wout^T                   (N=16)
//              tcgen05.ld d, dd/dx0, dd/dx1 (6 floats) ; det J, R, Euler update in fp32 registers
//   epilogue:  base log-prob (pdf mode), domain mapping + Jacobian, store wo / pdf
//
// The thread that issues a group's tcgen05.mma is one elected lane of the group itself (after a
// 128-thread named barrier), so there is no separate control warp competing for issue slots and no
// second mbarrier hop; completion comes back through tcgen05.commit -> mbarrier d_ready[g].
//
// Operands: A (activations; value + two tangent columns = three M=128 row blocks sharing B) lives in
// TMEM as fp16 (tcgen05.mma ".ts" form), written by the worker threads with tcgen05.st -- activations
// never touch shared memory or HBM.  B (weights) is resident in shared memory for the CTA's lifetime:
// the packed blob's fp16 image is ALREADY the UMMA canonical K-major layout, staged with one
// cp.async.bulk (TMA).  Every weight matrix is stored as fp16 hi + fp16 lo; the value path multiplies
// by both (weights effectively ~22 bits: weight rounding was the dominant fp16 error), the tangent path
// by hi only.  Accumulators are fp32 in TMEM.  x, det, R, pdf stay fp32 in registers for all T steps.
//
// tanh-form sigmoid: hidden-layer weights are pre-scaled by 1/2 (exact), so the MMA yields zh = z/2 and
// duh = du/2:   t = tanh(zh);  silu(z) = zh + zh t;  2 silu'(z) = (1 + t) + silu(z) (1 - t);
// u_out = 2 silu'(z) * duh  -> one MUFU op per activation and no rescaling.  The fp32 arithmetic is
// issued as packed f32x2 instructions (two activations per issue slot).
//
// TMEM map (512 columns allocated; per group 160 columns at g*160):
//   [  0, 96)  D_z | D_u | D_v   fp32 accumulators, 32 columns each (output round uses 16 of each)
//   [ 96,144)  A_h | A_u | A_v   fp16 operands, K=32 -> 16 columns each
//   [144,160)  A1                first-layer operand: cols 0..3 state (rewritten per step), 4..15 PE5(wi)
#include "common.cuh"
#include <cstdlib>

namespace bsdfdiff {

constexpr int kGroups = 3;
constexpr int kTcThreads = kGroups * 128;
constexpr int kTile = 128;
constexpr int kColsPerGroup = 160;
constexpr int kColD = 0, kColA = 96, kColA1 = 144;
constexpr int kTmemCols = 512;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ unsigned int g_tc_timeout_flag = 0;

// Bounded wait: a protocol bug must never hang the GPU.  try_wait suspends the warp in hardware (up to
// the hint) instead of spinning, so waiting warps do not steal issue slots from the computing ones.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(10000u) : "memory");
        if (done) return true;
    }
    atomicExch(&g_tc_timeout_flag, 1u);
    return false;
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {                   // exactly one lane of a converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void group_sync(int g) {            // named barrier 1+g over the group's 128 threads
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(128) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::f16 (fp16 operands, fp32 accumulate)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t x, y;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(taddr) : "memory");
    a = __uint_as_float(x); b = __uint_as_float(y);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float f16_round(float x) { return __half2float(__float2half_rn(x)); }

// packed fp32 pairs (sm_100 f32x2 ALU instructions: two lanes per issue slot)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint32_t pack_h2(f32x2 v) {
    float lo, hi; upk2(v, lo, hi); return pack_h2(lo, hi);
}
__device__ __forceinline__ float tanh_approx(float x) {
    float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x)); return t;
}

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B fp16, both K-major, M=128 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// activations on a pair of neurons.  Input zh = z/2 (packed pair).  Output h = silu(z), s2 = 2 silu'(z).
//   ACT 0: fp32 exp form (cross-check variant), ACT 1: tanh.approx.f32 + f32x2 arithmetic
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void silu_pair2(float zh0, float zh1, f32x2& h, f32x2& s2) {
    if (ACT == 0) {
        float hh[2], ss[2];
        const float zz[2] = {2.0f * zh0, 2.0f * zh1};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float s = __fdividef(1.0f, 1.0f + __expf(-zz[i]));
            hh[i] = zz[i] * s;
            ss[i] = 2.0f * s * fmaf(zz[i], 1.0f - s, 1.0f);
        }
        h = pk2(hh[0], hh[1]); s2 = pk2(ss[0], ss[1]);
    } else {
        const f32x2 one = pk2(1.0f, 1.0f);
        const f32x2 zh = pk2(zh0, zh1);
        const f32x2 t = pk2(tanh_approx(zh0), tanh_approx(zh1));
        h = fma2(zh, t, zh);                       // zh (1 + t)
        s2 = fma2(h, sub2(one, t), add2(one, t));  // (1 + t) + h (1 - t)
    }
}

// PE5(v) with two accurate sincosf and four double-angle steps (abs error ~1e-6, far below fp16 ulp)
__device__ __forceinline__ void pe5_fast(float v0, float v1, float* e) {
    float s0, c0, s1, c1;
    sincosf(v0, &s0, &c0);
    sincosf(v1, &s1, &c1);
    e[0] = v0; e[1] = v1;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        e[2 + 4 * k + 0] = s0; e[2 + 4 * k + 1] = s1; e[2 + 4 * k + 2] = c0; e[2 + 4 * k + 3] = c1;
        const float ns0 = 2.0f * s0 * c0, ns1 = 2.0f * s1 * c1;
        c0 = fmaf(c0, c0, -s0 * s0); c1 = fmaf(c1, c1, -s1 * s1);
        s0 = ns0; s1 = ns1;
    }
}

struct TcSmem {
    unsigned long long d_ready[kGroups];
    unsigned long long w_bar;
    uint32_t tmem_base;
    uint32_t pad[3];
    float base[kBaseFloats + 4];
    float4 aux[32];                         // per neuron: 0.5*W1[j,0], 0.5*W1[j,1], 0.5*W1[j,2], 0
    __align__(128) unsigned char w16[2 * (32 * 32 + 5 * 32 * 32 + 16 * 32) * 2];   // hi+lo images, <= 6 hidden layers
};

// Issue one round of MMAs for a group (executed by ONE thread).  Each layer's operand image = HI [N x 32]
// then LO [N x 32] (W = hi + lo): the value path accumulates A.hi^T + A.lo^T, tangents use hi only.
template <bool TANGENTS>
__device__ __forceinline__ void issue_round(int type, int NH, uint32_t tg, uint32_t w_base, uint32_t bar) {
    constexpr uint32_t idesc32 = make_idesc(32), idesc16 = make_idesc(16);
    if (type == 0) {                                // layer 1: D_z = A1 . W1^T  (K = 32 -> 2 x K16 per image)
        const uint64_t b = make_b_desc(w_base, 512, 128);
        mma_ts(tg + kColD, tg + kColA1, b, idesc32, 0u);
        mma_ts(tg + kColD, tg + kColA1 + 8, b + (1024 >> 4), idesc32, 1u);
        mma_ts(tg + kColD, tg + kColA1, b + (2048 >> 4), idesc32, 1u);
        mma_ts(tg + kColD, tg + kColA1 + 8, b + (3072 >> 4), idesc32, 1u);
    } else if (type < NH) {                         // hidden layer (type+1)
        const uint64_t b = make_b_desc(w_base + 4096u * type, 512, 128);
        mma_ts(tg + kColD, tg + kColA, b, idesc32, 0u);
        mma_ts(tg + kColD, tg + kColA + 8, b + (1024 >> 4), idesc32, 1u);
        mma_ts(tg + kColD, tg + kColA, b + (2048 >> 4), idesc32, 1u);
        mma_ts(tg + kColD, tg + kColA + 8, b + (3072 >> 4), idesc32, 1u);
        if (TANGENTS) {
#pragma unroll
            for (int c = 1; c < 3; ++c) {
                mma_ts(tg + kColD + 32 * c, tg + kColA + 16 * c, b, idesc32, 0u);
                mma_ts(tg + kColD + 32 * c, tg + kColA + 16 * c + 8, b + (1024 >> 4), idesc32, 1u);
            }
        }
    } else {                                        // output layer, N = 16
        const uint64_t b = make_b_desc(w_base + 4096u * NH, 256, 128);
        mma_ts(tg + kColD, tg + kColA, b, idesc16, 0u);
        mma_ts(tg + kColD, tg + kColA + 8, b + (512 >> 4), idesc16, 1u);
        mma_ts(tg + kColD, tg + kColA, b + (1024 >> 4), idesc16, 1u);
        mma_ts(tg + kColD, tg + kColA + 8, b + (1536 >> 4), idesc16, 1u);
        if (TANGENTS) {
#pragma unroll
            for (int c = 1; c < 3; ++c) {
                mma_ts(tg + kColD + 32 * c, tg + kColA + 16 * c, b, idesc16, 0u);
                mma_ts(tg + kColD + 32 * c, tg + kColA + 16 * c + 8, b + (512 >> 4), idesc16, 1u);
            }
        }
    }
    tc_commit(bar);
}

// Optional in-kernel phase timers (PROF instantiation only, selected with BSDFDIFF_TC_PROFILE=1): cycles per
// warp spent in  0 prologue/state-pack, 1 waiting for MMA results, 2 tcgen05.ld + activation math + st issue,
// 3 wait::st + group barrier, 4 MMA issue, 5 output round + epilogue, 6 total.
constexpr int kProfSlots = 8;
__device__ unsigned long long g_tc_prof[148 * 12 * kProfSlots];
#define PROF_T(slot)                                                     \
    do {                                                                 \
        if (PROF) {                                                      \
            const long long now_ = clock64();                            \
            prof[slot] += (unsigned long long)(now_ - tlast);            \
            tlast = now_;                                                \
        }                                                                \
    } while (0)

template <bool TANGENTS, int ACT, bool PROF>
__global__ void __launch_bounds__(kTcThreads, 1) flow_tc_kernel(const FlowParams P) {
    __shared__ TcSmem S;
    // warp index through a shuffle so the compiler can prove it warp-uniform: TMEM addresses and MMA
    // descriptors then live in uniform registers (no per-MMA R2UR/ELECT waterfall)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    unsigned long long prof[kProfSlots] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = PROF ? clock64() : 0;
    const long long tstart = tlast;
    const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(P.flow);
    const int NH = P.n_hidden;

    // ---- one-time setup ---------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int g = 0; g < kGroups; ++g) mbar_init(smem_u32(&S.d_ready[g]), 1);
        mbar_init(smem_u32(&S.w_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t f16_bytes = hdr->f16_bytes;
        mbar_expect_tx(smem_u32(&S.w_bar), f16_bytes);
        tma_bulk_g2s(smem_u32(S.w16), P.flow + hdr->off_f16, f16_bytes, smem_u32(&S.w_bar));
    }
    {
        const float* aux = reinterpret_cast<const float*>(P.flow + hdr->reserved[0]);
        if (threadIdx.x < 32)
            S.aux[threadIdx.x] = make_float4(aux[threadIdx.x], aux[32 + threadIdx.x], aux[64 + threadIdx.x], 0.0f);
        if (P.base) for (int i = threadIdx.x; i < kBaseFloats; i += kTcThreads) S.base[i] = P.base[i];
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&S.tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, S.tmem_base, 0);

    const long long n_tiles = (P.n + kTile - 1) / kTile;
    // tile list of this CTA: blockIdx.x, +gridDim.x, ...; group g takes every kGroups-th entry
    const long long my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    const int g = warp >> 2, q = warp & 3;
    const uint32_t tg_mma = tmem_base + g * kColsPerGroup;                   // lane field 0: MMA operand addresses
    const uint32_t tg = tg_mma + ((uint32_t)(q * 32) << 16);                 // this warp's 32-lane window
    const uint32_t bar_d = smem_u32(&S.d_ready[g]);
    const uint32_t w_base = smem_u32(S.w16);
    uint32_t pd = 0;
    bool ok = true;
    const float inv_t = (float)(1.0 / (double)P.T);
    const float sgn = (P.mode == kModePdf) ? -1.0f : 1.0f;
    const float step = sgn * inv_t;
    if (q == 0) ok = mbar_wait(smem_u32(&S.w_bar), 0);                      // weights have landed in smem

    for (long long k = g; k < my_tiles && ok; k += kGroups) {
        const long long tile = blockIdx.x + k * gridDim.x;
        const long long i_raw = tile * kTile + q * 32 + lane;
        const bool valid = i_raw < P.n;
        const long long i = valid ? i_raw : (P.n - 1);      // tail rows recompute the last query, never store

        float w0, w1, wiz;
        load_wi(P, i, w0, w1, wiz);
        float bp[4] = {0.f, 0.f, 0.f, 0.f};
        {
            float e[kPE5];
            pe5_fast(w0, w1, e);
            uint32_t c[12];
#pragma unroll
            for (int j = 0; j < 11; ++j) c[j] = pack_h2(e[2 * j], e[2 * j + 1]);
            c[11] = 0u;
            tmem_st8(tg + kColA1 + 4, c);
            tmem_st4(tg + kColA1 + 12, c[8], c[9], c[10], c[11]);
            if (P.base) {       // base net shares the first three PE frequencies (PE3 is a prefix of PE5)
                const float* b = S.base;
                bp[0] = b[304]; bp[1] = b[305]; bp[2] = b[306]; bp[3] = b[307];
#pragma unroll 4
                for (int j = 0; j < 16; ++j) {
                    float z = b[224 + j];
#pragma unroll
                    for (int kk = 0; kk < kPE3; ++kk) z = fmaf(e[kk], b[j * kPE3 + kk], z);
                    const float h = z * __fdividef(1.0f, 1.0f + __expf(-z));
                    bp[0] = fmaf(h, b[240 + j], bp[0]); bp[1] = fmaf(h, b[256 + j], bp[1]);
                    bp[2] = fmaf(h, b[272 + j], bp[2]); bp[3] = fmaf(h, b[288 + j], bp[3]);
                }
            }
        }

        float x0, x1, R = 1.0f, p0 = 1.0f;
        float wox = 0.0f, woy = 0.0f, woz = 1.0f, theta_o = 0.0f;
        if (P.mode == kModePdf) {
            load_wo(P, i, x0, x1, wox, woy, woz);
            theta_o = x0;
        } else {
            if (P.x0) {
                const float2 t = reinterpret_cast<const float2*>(P.x0)[i];
                x0 = t.x; x1 = t.y;
            } else {
                base_draw(P.domain, bp, P.seed, P.offset, P.first_index + i, x0, x1);
            }
            if (P.out_x0 && valid) reinterpret_cast<float2*>(P.out_x0)[i] = make_float2(x0, x1);
            if (P.mode == kModeSample) p0 = expf(base_logprob(P.domain, bp, x0, x1));
        }

        for (int t = 0; t < P.T && ok; ++t) {
            const float tf = (float)t / (float)P.T;
            const float alpha = (P.mode == kModePdf) ? 1.0f - tf : tf;
            // ---- state -> A1 columns 0..3 (hi parts, then lo parts) ----
            float sphi = 0.0f, cphi = 1.0f;
            {
                float s0, s1, s2v, s3;
                if (P.domain == kDisk) { s0 = x0; s1 = x1; s2v = alpha; s3 = 0.0f; }
                else { __sincosf(x1, &sphi, &cphi); s0 = x0; s1 = sphi; s2v = cphi; s3 = alpha; }
                const float h0 = f16_round(s0), h1 = f16_round(s1), h2 = f16_round(s2v), h3 = f16_round(s3);
                uint32_t c0, c1, c2, c3;
                if (P.domain == kDisk) {
                    // k: x0_hi x1_hi | a_hi a_lo | x0_lo x1_lo | 0 0
                    c0 = pack_h2(h0, h1); c1 = pack_h2(h2, s2v - h2); c2 = pack_h2(s0 - h0, s1 - h1); c3 = 0u;
                } else {
                    // k: th_hi sin_hi | cos_hi a_hi | th_lo sin_lo | cos_lo a_lo
                    c0 = pack_h2(h0, h1); c1 = pack_h2(h2, h3); c2 = pack_h2(s0 - h0, s1 - h1);
                    c3 = pack_h2(s2v - h2, s3 - h3);
                }
                tmem_st4(tg + kColA1, c0, c1, c2, c3);
            }
            PROF_T(0);
            tc_wait_st();
            tc_fence_before();
            group_sync(g);
            PROF_T(3);
            if (q == 0 && elect_one()) { tc_fence_after(); issue_round<TANGENTS>(0, NH, tg_mma, w_base, bar_d); }
            __syncwarp();
            PROF_T(4);

            // ---- round 0: first layer ----
            ok = mbar_wait(bar_d, pd); pd ^= 1u;
            __syncwarp();
            PROF_T(1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float z[16];
                tmem_ld16(tg + kColD + 16 * half, z);
                tc_wait_ld();
                uint32_t ph[8], pu[8], pv[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    f32x2 h, s2;
                    silu_pair2<ACT>(z[j], z[j + 1], h, s2);
                    ph[j >> 1] = pack_h2(h);
                    if (TANGENTS) {
                        const float4 a0 = S.aux[16 * half + j], a1 = S.aux[16 * half + j + 1];
                        f32x2 su, sv;
                        su = pk2(a0.x, a1.x);
                        if (P.domain == kDisk) sv = pk2(a0.y, a1.y);
                        else sv = pk2(fmaf(cphi, a0.y, -sphi * a0.z), fmaf(cphi, a1.y, -sphi * a1.z));
                        pu[j >> 1] = pack_h2(mul2(s2, su));
                        pv[j >> 1] = pack_h2(mul2(s2, sv));
                    }
                }
                tmem_st8(tg + kColA + 8 * half, ph);
                if (TANGENTS) { tmem_st8(tg + kColA + 16 + 8 * half, pu); tmem_st8(tg + kColA + 32 + 8 * half, pv); }
            }
            PROF_T(2);
            tc_wait_st();
            tc_fence_before();
            group_sync(g);
            PROF_T(3);
            if (q == 0 && elect_one()) { tc_fence_after(); issue_round<TANGENTS>(1, NH, tg_mma, w_base, bar_d); }
            __syncwarp();
            PROF_T(4);

            // ---- hidden rounds ----
            for (int l = 1; l < NH && ok; ++l) {
                ok = mbar_wait(bar_d, pd); pd ^= 1u;
                __syncwarp();
                PROF_T(1);
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float z[16], du[TANGENTS ? 16 : 1], dv[TANGENTS ? 16 : 1];
                    tmem_ld16(tg + kColD + 16 * half, z);
                    if (TANGENTS) { tmem_ld16(tg + kColD + 32 + 16 * half, du); tmem_ld16(tg + kColD + 64 + 16 * half, dv); }
                    tc_wait_ld();
                    uint32_t ph[8], pu[8], pv[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        f32x2 h, s2;
                        silu_pair2<ACT>(z[j], z[j + 1], h, s2);
                        ph[j >> 1] = pack_h2(h);
                        if (TANGENTS) {
                            pu[j >> 1] = pack_h2(mul2(s2, pk2(du[j], du[j + 1])));
                            pv[j >> 1] = pack_h2(mul2(s2, pk2(dv[j], dv[j + 1])));
                        }
                    }
                    tmem_st8(tg + kColA + 8 * half, ph);
                    if (TANGENTS) { tmem_st8(tg + kColA + 16 + 8 * half, pu); tmem_st8(tg + kColA + 32 + 8 * half, pv); }
                }
                PROF_T(2);
                tc_wait_st();
                tc_fence_before();
                group_sync(g);
                PROF_T(3);
                if (q == 0 && elect_one()) { tc_fence_after(); issue_round<TANGENTS>(l + 1, NH, tg_mma, w_base, bar_d); }
                __syncwarp();
                PROF_T(4);
            }

            // ---- output round: d, dd/dx0, dd/dx1 ----
            ok = ok && mbar_wait(bar_d, pd); pd ^= 1u;
            __syncwarp();
            PROF_T(1);
            tc_fence_after();
            float d0, d1, du0 = 0.f, du1 = 0.f, dv0 = 0.f, dv1 = 0.f;
            tmem_ld2(tg + kColD, d0, d1);
            if (TANGENTS) { tmem_ld2(tg + kColD + 32, du0, du1); tmem_ld2(tg + kColD + 64, dv0, dv1); }
            tc_wait_ld();
            if (TANGENTS) {
                const float j00 = fmaf(step, du0, 1.0f), j01 = step * dv0;
                const float j10 = step * du1, j11 = fmaf(step, dv1, 1.0f);
                const float det = j00 * j11 - j01 * j10;
                R = (P.mode == kModePdf) ? R * det : R / det;
            }
            x0 = fmaf(step, d0, x0);
            x1 = fmaf(step, d1, x1);
            PROF_T(5);
        }

        if (valid && ok) {
            if (P.mode == kModeSample) {
                store_sample(P, i, x0, x1, p0 * R);
            } else if (P.mode == kModePdf) {
                store_pdf(P, i, expf(base_logprob(P.domain, bp, x0, x1)) * R, wiz, wox, woy, woz, theta_o);
            } else {
                reinterpret_cast<float2*>(P.out_dir)[i] = make_float2(x0, x1);
            }
        }
        PROF_T(5);
    }
    if (PROF && lane == 0 && blockIdx.x < 148) {
        prof[6] = (unsigned long long)(clock64() - tstart);
        for (int s = 0; s < kProfSlots; ++s) g_tc_prof[(blockIdx.x * 12 + warp) * kProfSlots + s] = prof[s];
    }

    // ---- teardown ---------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                     : "memory");
    }
}

template <bool TANGENTS, int ACT>
static int launch_tc_t(const FlowParams& P, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (P.n + kTile - 1) / kTile;
    long long grid = sms;
    if (grid > tiles) grid = tiles;
    if (grid < 1) return 0;
    static const bool profile = (getenv("BSDFDIFF_TC_PROFILE") != nullptr);    // bring-up / tuning aid only
    if (profile && TANGENTS && ACT == 1)
        flow_tc_kernel<TANGENTS, ACT, true><<<(unsigned)grid, kTcThreads, 0, stream>>>(P);
    else
        flow_tc_kernel<TANGENTS, ACT, false><<<(unsigned)grid, kTcThreads, 0, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// variant 1 = tc16 (tanh.approx activation), 2 = tc16 with fp32 exp-form activation (cross-check)
int launch_tc(const FlowParams& P, cudaStream_t stream, int variant) {
    if (P.hidden != 32 || P.n_hidden < 2 || P.n_hidden > 6) return -2;      // 64-wide nets: CUDA-core path
    const bool tang = (P.mode != kModeForward);
    if (variant == 2) return tang ? launch_tc_t<true, 0>(P, stream) : launch_tc_t<false, 0>(P, stream);
    return tang ? launch_tc_t<true, 1>(P, stream) : launch_tc_t<false, 1>(P, stream);
}

unsigned int tc_timeout_flag() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_tc_timeout_flag, sizeof(v));
    return v;
}

// copies the phase timers of the last profiled launch: [148 CTAs][12 warps][8 slots] cycles (synchronises)
int tc_profile_fetch(unsigned long long* out, int max_elems) {
    const int n = 148 * 12 * kProfSlots;
    if (max_elems < n) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(unsigned long long) * n) == cudaSuccess ? n : -3;
}

}  // namespace bsdfdiff
