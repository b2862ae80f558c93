// Micro-benchmarks behind the kernel design decisions in DESIGN.md (run on one B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench profiles/microbench/ubench.cu && ./gpurun_out/ubench
// Measures, per SM, with clock64 inside the kernel:
//   1. tcgen05.ld / tcgen05.st throughput vs number of warps
//   2. MUFU tanh (f32, f16x2), f32x2 FMA, f16x2 pack throughput vs number of warps
//   3. tcgen05.mma round trip: issue k MMAs (M128 N32 K16, A in TMEM) + commit -> mbarrier wake-up
//   4. the full per-round chain of the flow kernel (st -> fence -> bar -> issue -> commit -> wait -> ld)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (int it = 0; it < (1 << 16) && !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    }
    return done != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// K MMAs (M128 N32 K16, A in TMEM) with compile-time offsets from warp-uniform bases, then commit
template <int K>
__device__ __forceinline__ void issue_k(uint32_t d0, uint32_t a0, uint64_t b, uint32_t bar) {
#pragma unroll
    for (int k = 0; k < K; ++k)
        mma_ts(d0 + 32 * (k % 3), a0 + 16 * (k % 3) + 8 * ((k / 3) & 1), b, make_idesc(32), (k >= 3) ? 1u : 0u);
    tc_commit(bar);
}
__device__ __forceinline__ void issue_rt(int K, uint32_t d0, uint32_t a0, uint64_t b, uint32_t bar) {
    switch (K) {
        case 1: issue_k<1>(d0, a0, b, bar); break;
        case 2: issue_k<2>(d0, a0, b, bar); break;
        case 3: issue_k<3>(d0, a0, b, bar); break;
        case 6: issue_k<6>(d0, a0, b, bar); break;
        case 8: issue_k<8>(d0, a0, b, bar); break;
        default: issue_k<12>(d0, a0, b, bar); break;
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// pipe-sharing probe: NT tanh.f32, NC cvt.f16x2.f32, NF fma.f32x2, NH fma.f16x2, NP prmt per inner iteration,
// all independent (8 chains each), compile-time mix -> no branches in the loop
template <int NT, int NC, int NF, int NH, int NP>
__device__ __forceinline__ long long pipe_probe(int lane, uint32_t& sink) {
    float x[8]; uint32_t y[8]; unsigned long long w2[8]; uint32_t hh[8]; uint32_t pp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = 0.01f * (lane + i); y[i] = 0x3e000000u + (lane << 8) + i; w2[i] = 0x3f8000003f800000ull + lane + i; hh[i] = 0x3c003c00u + i; pp[i] = lane * i; }
    const unsigned long long m = 0x3f7fff003f7fff00ull;
    const uint32_t hm = 0x3bff3bffu;
    float p0 = 0.3f + lane, p1 = 0.7f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NT) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            if (i < NC) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y[i]) : "f"(__uint_as_float(y[i])), "f"(p1));
            if (i < NF) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w2[i]) : "l"(m));
            if (i < NH) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(hh[i]) : "r"(hm));
            if (i < NP) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(pp[i]) : "r"(hm));
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) sink += __float_as_uint(x[i]) + y[i] + (uint32_t)w2[i] + hh[i] + pp[i];
    return t1 - t0;
}

// activation math on registers only (no TMEM): 8 independent neuron pairs per iteration.
//   F32: 2 tanh.f32 + fma2, add2, sub2, fma2, mul2 x2 + 3 cvt.f16x2 per pair   (the tc16 inner loop)
//   H2 : 1 tanh.f16x2 + 6 half2 ops per pair                                  (fp16-accumulator variant)
template <bool H2>
__device__ __forceinline__ long long math_probe(int lane, uint32_t& sink) {
    long long t0, t1;
    if (!H2) {
        float z[16], du[16], dv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { z[i] = 0.01f * (lane + i); du[i] = 0.5f + i; dv[i] = 0.25f * i; }
        uint32_t acc = 0;
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 256; ++it) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float t0f, t1f;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0f) : "f"(z[j]));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1f) : "f"(z[j + 1]));
                unsigned long long zz, tt, one, h, s2, a, b, u, v, duu, dvv;
                asm("mov.b64 %0, {%1, %2};" : "=l"(zz) : "f"(z[j]), "f"(z[j + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(tt) : "f"(t0f), "f"(t1f));
                asm("mov.b64 %0, {%1, %2};" : "=l"(one) : "f"(1.0f), "f"(1.0f));
                asm("mov.b64 %0, {%1, %2};" : "=l"(duu) : "f"(du[j]), "f"(du[j + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(dvv) : "f"(dv[j]), "f"(dv[j + 1]));
                asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(h) : "l"(zz), "l"(tt));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(one), "l"(tt));
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(one), "l"(tt));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(s2) : "l"(h), "l"(a), "l"(b));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(s2), "l"(duu));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(s2), "l"(dvv));
                float lo, hi; uint32_t ph, pu, pv;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(h));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph) : "f"(hi), "f"(lo));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pu) : "f"(hi), "f"(lo));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pv) : "f"(hi), "f"(lo));
                const uint32_t m = (ph ^ pu ^ pv) & 1u;                  // loop-carried dependency (2 ALU ops + 1)
                z[j] = __uint_as_float(__float_as_uint(z[j]) ^ m);
                acc += m;
            }
        }
        t1 = clock64();
        sink += acc;
    } else {
        uint32_t z[8], du[8], dv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { z[i] = 0x30003000u + lane + i; du[i] = 0x38003800u + i; dv[i] = 0x34003400u + i; }
        uint32_t acc = 0;
        const uint32_t one = 0x3c003c00u;
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 256; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t t, h, a, b, s2, u, v;
                asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(z[j]));
                asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(h) : "r"(z[j]), "r"(t));
                asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(one), "r"(t));
                asm("add.rn.f16x2 %0, %1, %2;" : "=r"(b) : "r"(one), "r"(t));
                asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(s2) : "r"(h), "r"(a), "r"(b));
                asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(u) : "r"(s2), "r"(du[j]));
                asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(v) : "r"(s2), "r"(dv[j]));
                const uint32_t m = (h ^ u ^ v) & 1u;
                z[j] ^= m;
                acc += m;
            }
        }
        t1 = clock64();
        sink += acc;
    }
    return t1 - t0;
}

// Alternative formulations of the activation inner loop on registers only (8 neuron pairs per iteration; the only
// extra work is 2 LOP3 per pair that feed the outputs back into z so nothing is hoisted):
//   0 f32x2 (shipping)   1 scalar fp32   2 f32 tanh + f32x2 h, half2 s2 and tangent products
//   3 f32x2 h and s2, tangent products in half2 (du, dv as packed halves: fp16 tangent accumulators)
//   4 all half2 (fp16 accumulators everywhere)
template <int V>
__device__ __forceinline__ long long form_probe(int lane, uint32_t& sink) {
    float z[16], du[16], dv[16];
    uint32_t duh[8], dvh[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { z[i] = 0.01f * (lane + i); du[i] = 0.5f + i; dv[i] = 0.25f * i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { duh[i] = 0x38003800u + i; dvh[i] = 0x34003400u + i; }
    const uint32_t oneh = 0x3c003c00u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            uint32_t ph, pu, pv;
            if (V == 4) {
                uint32_t zz = __float_as_uint(z[j]), t, h, a, b, s2;
                asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(zz));
                asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(h) : "r"(zz), "r"(t));
                asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(oneh), "r"(t));
                asm("add.rn.f16x2 %0, %1, %2;" : "=r"(b) : "r"(oneh), "r"(t));
                asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(s2) : "r"(h), "r"(a), "r"(b));
                asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pu) : "r"(s2), "r"(duh[j >> 1]));
                asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pv) : "r"(s2), "r"(dvh[j >> 1]));
                ph = h;
            } else {
                float t0f, t1f;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0f) : "f"(z[j]));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1f) : "f"(z[j + 1]));
                if (V == 1) {
                    const float h0 = fmaf(z[j], t0f, z[j]), h1 = fmaf(z[j + 1], t1f, z[j + 1]);
                    const float s0 = fmaf(h0, 1.0f - t0f, 1.0f + t0f), s1 = fmaf(h1, 1.0f - t1f, 1.0f + t1f);
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph) : "f"(h1), "f"(h0));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pu) : "f"(s1 * du[j + 1]), "f"(s0 * du[j]));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pv) : "f"(s1 * dv[j + 1]), "f"(s0 * dv[j]));
                } else {
                    unsigned long long zz, tt, h;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(zz) : "f"(z[j]), "f"(z[j + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(tt) : "f"(t0f), "f"(t1f));
                    asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(h) : "l"(zz), "l"(tt));
                    float lo, hi;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(h));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph) : "f"(hi), "f"(lo));
                    if (V == 2) {
                        uint32_t th, a, b, s2;
                        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(th) : "f"(t1f), "f"(t0f));
                        asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(oneh), "r"(th));
                        asm("add.rn.f16x2 %0, %1, %2;" : "=r"(b) : "r"(oneh), "r"(th));
                        asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(s2) : "r"(ph), "r"(a), "r"(b));
                        asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pu) : "r"(s2), "r"(duh[j >> 1]));
                        asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pv) : "r"(s2), "r"(dvh[j >> 1]));
                    } else {
                        unsigned long long one, a, b, s2;
                        asm("mov.b64 %0, {%1, %2};" : "=l"(one) : "f"(1.0f), "f"(1.0f));
                        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(one), "l"(tt));
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(one), "l"(tt));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(s2) : "l"(h), "l"(a), "l"(b));
                        if (V == 0) {
                            unsigned long long duu, dvv, u, v;
                            asm("mov.b64 %0, {%1, %2};" : "=l"(duu) : "f"(du[j]), "f"(du[j + 1]));
                            asm("mov.b64 %0, {%1, %2};" : "=l"(dvv) : "f"(dv[j]), "f"(dv[j + 1]));
                            asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(s2), "l"(duu));
                            asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(s2), "l"(dvv));
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u));
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pu) : "f"(hi), "f"(lo));
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pv) : "f"(hi), "f"(lo));
                        } else {   // V == 3
                            uint32_t s2h;
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s2));
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(s2h) : "f"(hi), "f"(lo));
                            asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pu) : "r"(s2h), "r"(duh[j >> 1]));
                            asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pv) : "r"(s2h), "r"(dvh[j >> 1]));
                        }
                    }
                }
            }
            uint32_t m;
            asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(m) : "r"(ph), "r"(pu), "r"(pv));                      // xor3
            asm("lop3.b32 %0, %1, %2, 1, 0x6a;" : "=r"(m) : "r"(__float_as_uint(z[j])), "r"(m));              // z ^ (m & 1)
            z[j] = __uint_as_float(m);
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) sink += __float_as_uint(z[i]);
    return t1 - t0;
}

// MUFU / FP32 overlap at realistic ratios: NT independent tanh.f32 chains and NS independent scalar FFMA chains
template <int NT, int NS>
__device__ __forceinline__ long long overlap_probe(int lane, uint32_t& sink) {
    float x[NT > 0 ? NT : 1], y[NS > 0 ? NS : 1];
#pragma unroll
    for (int i = 0; i < NT; ++i) x[i] = 0.01f * (lane + i);
#pragma unroll
    for (int i = 0; i < NS; ++i) y[i] = 1.0f + 0.001f * (lane + i);
    const float m = 0.999f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < (NT > NS ? NT : NS); ++i) {
            if (i < NT) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            if (i < NS) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(y[i]) : "f"(m));
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < NT; ++i) sink += __float_as_uint(x[i]);
#pragma unroll
    for (int i = 0; i < NS; ++i) sink += __float_as_uint(y[i]);
    return t1 - t0;
}

struct Res { long long cyc[16]; };

// test ids
enum { T_LD16 = 0, T_LD32, T_ST8, T_TANH32, T_TANH16, T_FMA2, T_PACK, T_MIX, T_MMA, T_CHAIN, T_CHAIN2,
       T_TANH_PACK, T_TANH_FMA2, T_PACK_FMA2, T_HFMA2, T_PRMT, T_TANH_HFMA2, T_PACK_HFMA2 };

__global__ void __launch_bounds__(512, 1) ubench(int test, int nwarps, int kparam, Res* out) {
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned long long bars[8];
    __shared__ __align__(128) unsigned char wsm[8192];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(wsm)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        for (int i = 4; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 4);
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
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    const uint32_t tw = tb + ((uint32_t)((warp & 3) * 32) << 16);
    // zero TMEM so tanh etc. see finite data
    {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (warp < 4) for (int c = 0; c < 512; c += 8) tmem_st8(tw + c, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    long long t0 = 0, t1 = 0;
    const int ITERS = 256;
    uint32_t sink = 0;
    if (warp < nwarps) {
        const uint32_t colbase = (uint32_t)((warp >> 2) * 128) & 511u;     // warps sharing a quadrant use other columns
        if (test == T_LD16) {
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                uint32_t a[16], b[16], c[16], d[16];
                tmem_ld16(tw + colbase + 0, a); tmem_ld16(tw + colbase + 16, b);
                tmem_ld16(tw + colbase + 32, c); tmem_ld16(tw + colbase + 48, d);
                tc_wait_ld();
                sink += a[0] + b[1] + c[2] + d[3];
            }
            t1 = clock64();
        } else if (test == T_LD32) {
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                uint32_t a[32], b[32];
                tmem_ld32(tw + colbase + 0, a); tmem_ld32(tw + colbase + 32, b);
                tc_wait_ld();
                sink += a[0] + b[1] + a[31] + b[17];
            }
            t1 = clock64();
        } else if (test == T_ST8) {
            uint32_t v[8] = {1, 2, 3, 4, 5, 6, 7, (uint32_t)lane};
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                tmem_st8(tw + colbase + 0, v); tmem_st8(tw + colbase + 8, v); tmem_st8(tw + colbase + 16, v); tmem_st8(tw + colbase + 24, v);
                tmem_st8(tw + colbase + 32, v); tmem_st8(tw + colbase + 40, v); tmem_st8(tw + colbase + 48, v); tmem_st8(tw + colbase + 56, v);
                tc_wait_st();
            }
            t1 = clock64();
        } else if (test == T_TANH32) {
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = 0.01f * (lane + i);
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            }
            t1 = clock64();
#pragma unroll
            for (int i = 0; i < 16; ++i) sink += __float_as_uint(x[i]);
        } else if (test == T_TANH16) {
            uint32_t x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = 0x30003000u + lane + i;
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x[i]));
            }
            t1 = clock64();
#pragma unroll
            for (int i = 0; i < 16; ++i) sink += x[i];
        } else if (test == T_FMA2) {
            unsigned long long x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = 0x3f8000003f800000ull + lane + i;
            const unsigned long long m = 0x3f7fff003f7fff00ull;
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[i]) : "l"(m));
            }
            t1 = clock64();
#pragma unroll
            for (int i = 0; i < 16; ++i) sink += (uint32_t)x[i];
        } else if (test == T_PACK) {
            float x[16]; uint32_t y[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = 0.01f * (lane + i); y[i] = 0; }
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y[i]) : "f"(x[i]), "f"(x[(i + 1) & 15]));
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(y[i] | 0x30000000u);
            }
            t1 = clock64();
#pragma unroll
            for (int i = 0; i < 16; ++i) sink += y[i];
        } else if (test == T_MIX) {
            // the activation inner loop: 16 z + 16 du + 16 dv from TMEM -> tanh/f32x2 math -> 3 x st8
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                uint32_t z[16], du[16], dv[16], ph[8], pu[8], pv[8];
                tmem_ld16(tw + colbase + 0, z); tmem_ld16(tw + colbase + 32, du); tmem_ld16(tw + colbase + 64, dv);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float z0 = __uint_as_float(z[j]), z1 = __uint_as_float(z[j + 1]), t0f, t1f;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0f) : "f"(z0));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1f) : "f"(z1));
                    unsigned long long zz, tt, one, h, s2, a, b, u, v, duu, dvv;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(zz) : "f"(z0), "f"(z1));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(tt) : "f"(t0f), "f"(t1f));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(one) : "f"(1.0f), "f"(1.0f));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(duu) : "r"(du[j]), "r"(du[j + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(dvv) : "r"(dv[j]), "r"(dv[j + 1]));
                    asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(h) : "l"(zz), "l"(tt));
                    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(one), "l"(tt));
                    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(one), "l"(tt));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(s2) : "l"(h), "l"(a), "l"(b));
                    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(s2), "l"(duu));
                    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(s2), "l"(dvv));
                    float lo, hi;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(h));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph[j >> 1]) : "f"(hi), "f"(lo));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pu[j >> 1]) : "f"(hi), "f"(lo));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pv[j >> 1]) : "f"(hi), "f"(lo));
                }
                tmem_st8(tw + colbase + 96, ph); tmem_st8(tw + colbase + 104, pu); tmem_st8(tw + colbase + 112, pv);
            }
            tc_wait_st();
            t1 = clock64();
        }
    }
    if (test == T_MMA) {
        // warp 0 (uniform), one elected lane: issue kparam MMAs + commit, wait for the barrier
        if (warp == 0) {
            const uint64_t b = make_b_desc(smem_u32(wsm), 512, 128);
            uint32_t par = 0;
            long long acc_issue = 0, acc_total = 0;
            for (int rep = 0; rep < 64; ++rep) {
                tc_fence_after();
                const long long a0 = clock64();
                if (elect_one()) issue_rt(kparam, tb, tb + 256, b, smem_u32(&bars[0]));
                __syncwarp();
                const long long a1 = clock64();
                mbar_wait(smem_u32(&bars[0]), par); par ^= 1u;
                const long long a2 = clock64();
                acc_issue += a1 - a0; acc_total += a2 - a0;
            }
            if (lane == 0) { out[blockIdx.x].cyc[0] = acc_issue / 64; out[blockIdx.x].cyc[1] = acc_total / 64; }
        }
    } else if (test == T_CHAIN) {
        // per-round chain, group-local issue: st -> wait::st -> fence -> bar.sync(128) -> elected lane issues -> wait -> ld
        const int g = warp >> 2, q = warp & 3;
        if (warp < nwarps) {
            const uint32_t tg_mma = tb + g * 160, tg = tg_mma + ((uint32_t)(q * 32) << 16);
            const uint64_t b = make_b_desc(smem_u32(wsm), 512, 128);
            uint32_t par = 0;
            uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                tmem_st8(tg + 96, v); tmem_st8(tg + 104, v); tmem_st8(tg + 112, v);
                tmem_st8(tg + 120, v); tmem_st8(tg + 128, v); tmem_st8(tg + 136, v);
                tc_wait_st();
                tc_fence_before();
                asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(128) : "memory");
                if (q == 0) {
                    if (elect_one()) { tc_fence_after(); issue_rt(kparam, tg_mma, tg_mma + 96, b, smem_u32(&bars[g])); }
                    __syncwarp();
                }
                mbar_wait(smem_u32(&bars[g]), par); par ^= 1u;
                tc_fence_after();
                uint32_t a[16], c[16], d[16];
                tmem_ld16(tg + 0, a); tmem_ld16(tg + 32, c); tmem_ld16(tg + 64, d);
                tc_wait_ld();
                v[0] = a[0] & c[0] & d[0] & 0u;
            }
            t1 = clock64();
        }
    } else if (test == T_CHAIN2) {
        // dedicated issuer warps 12..14 (one per group); workers arrive on an mbarrier (count 4), no bar.sync
        const int ngroups = nwarps / 4;
        if (warp < nwarps) {
            const int g = warp >> 2, q = warp & 3;
            const uint32_t tg_mma = tb + g * 160, tg = tg_mma + ((uint32_t)(q * 32) << 16);
            uint32_t par = 0;
            uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                tmem_st8(tg + 96, v); tmem_st8(tg + 104, v); tmem_st8(tg + 112, v);
                tmem_st8(tg + 120, v); tmem_st8(tg + 128, v); tmem_st8(tg + 136, v);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[4 + g]));
                mbar_wait(smem_u32(&bars[g]), par); par ^= 1u;
                tc_fence_after();
                uint32_t a[16], c[16], d[16];
                tmem_ld16(tg + 0, a); tmem_ld16(tg + 32, c); tmem_ld16(tg + 64, d);
                tc_wait_ld();
                v[0] = a[0] & c[0] & d[0] & 0u;
            }
            t1 = clock64();
        } else if (warp >= 12 && warp < 12 + ngroups) {
            const int g = warp - 12;
            const uint32_t tg_mma = tb + g * 160;
            const uint64_t b = make_b_desc(smem_u32(wsm), 512, 128);
            uint32_t par = 0;
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                mbar_wait(smem_u32(&bars[4 + g]), par); par ^= 1u;
                tc_fence_after();
                if (elect_one()) issue_rt(kparam, tg_mma, tg_mma + 96, b, smem_u32(&bars[g]));
                __syncwarp();
            }
        }
    } else if (warp < nwarps && test >= T_TANH_PACK) {
        t0 = 0;
        switch (kparam) {
            case 0: t1 = pipe_probe<8, 0, 0, 0, 0>(lane, sink); break;   // tanh
            case 1: t1 = pipe_probe<0, 8, 0, 0, 0>(lane, sink); break;   // cvt
            case 2: t1 = pipe_probe<0, 0, 8, 0, 0>(lane, sink); break;   // fma2
            case 3: t1 = pipe_probe<0, 0, 0, 8, 0>(lane, sink); break;   // hfma2
            case 4: t1 = pipe_probe<0, 0, 0, 0, 8>(lane, sink); break;   // prmt
            case 5: t1 = pipe_probe<8, 8, 0, 0, 0>(lane, sink); break;   // tanh + cvt
            case 6: t1 = pipe_probe<8, 0, 8, 0, 0>(lane, sink); break;   // tanh + fma2
            case 7: t1 = pipe_probe<0, 8, 8, 0, 0>(lane, sink); break;   // cvt + fma2
            case 8: t1 = pipe_probe<8, 0, 0, 8, 0>(lane, sink); break;   // tanh + hfma2
            case 9: t1 = pipe_probe<0, 8, 0, 8, 0>(lane, sink); break;   // cvt + hfma2
            case 10: t1 = pipe_probe<0, 0, 8, 8, 0>(lane, sink); break;  // fma2 + hfma2
            case 11: t1 = pipe_probe<8, 8, 8, 0, 0>(lane, sink); break;  // tanh + cvt + fma2
            case 12: t1 = pipe_probe<4, 6, 8, 0, 0>(lane, sink); break;  // the activation mix: 2 tanh : 3 cvt : 4(of 6) fma2
            case 13: t1 = pipe_probe<0, 8, 0, 0, 8>(lane, sink); break;  // cvt + prmt
            case 14: t1 = pipe_probe<8, 0, 0, 0, 8>(lane, sink); break;  // tanh + prmt
            case 15: t1 = math_probe<false>(lane, sink); break;
            case 16: t1 = math_probe<true>(lane, sink); break;
            case 17: t1 = form_probe<0>(lane, sink); break;
            case 18: t1 = form_probe<1>(lane, sink); break;
            case 19: t1 = form_probe<2>(lane, sink); break;
            case 20: t1 = form_probe<3>(lane, sink); break;
            case 21: t1 = form_probe<4>(lane, sink); break;
            case 22: t1 = overlap_probe<8, 0>(lane, sink); break;
            case 23: t1 = overlap_probe<0, 48>(lane, sink); break;
            case 24: t1 = overlap_probe<8, 16>(lane, sink); break;
            case 25: t1 = overlap_probe<8, 32>(lane, sink); break;
            case 26: t1 = overlap_probe<8, 48>(lane, sink); break;
            default: t1 = overlap_probe<8, 64>(lane, sink); break;
        }
    }
    if (lane == 0 && warp < 16 && test != T_MMA) out[blockIdx.x].cyc[warp] = (t1 - t0);
    if (sink == 0x12345678u) out[blockIdx.x].cyc[15] = sink;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    }
}

static int run(const char* name, int test, int nwarps, int k, int grid, double unit_per_iter_per_warp, const char* unit) {
    Res* d; CK(cudaMalloc(&d, sizeof(Res) * grid)); CK(cudaMemset(d, 0, sizeof(Res) * grid));
    ubench<<<grid, 512>>>(test, nwarps, k, d);
    CK(cudaDeviceSynchronize());
    Res* h = new Res[grid];
    CK(cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost));
    if (test == T_MMA) {
        printf("%-10s k=%2d grid=%3d: issue %lld cyc, issue->wake %lld cyc\n", name, k, grid, h[0].cyc[0], h[0].cyc[1]);
    } else {
        long long mx = 0;
        for (int w = 0; w < nwarps; ++w) mx = h[0].cyc[w] > mx ? h[0].cyc[w] : mx;
        const double per_iter = (double)mx / 256.0;
        printf("%-18s warps=%2d k=%2d grid=%3d: %8.1f cyc/iter/warp  -> %8.2f %s per clk per SM\n", name, nwarps, k, grid, per_iter,
               unit_per_iter_per_warp * nwarps / per_iter, unit);
    }
    delete[] h; cudaFree(d);
    return 0;
}

// ---- functional probe: fp16 accumulators (idesc D-format f16) and tcgen05.ld .pack::16b -----------------
__global__ void __launch_bounds__(128, 1) f16acc_probe(uint32_t* out) {
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned long long bar;
    __shared__ __align__(128) __half wsm[32 * 16];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 16; i += 128) {
        const int n = i / 16, k = i % 16;                   // B[n][k]: n<16 -> 1 at k==n ; n>=16 -> 2 at k==n-16
        const float v = (n < 16) ? (k == n ? 1.0f : 0.0f) : (k == n - 16 ? 2.0f : 0.0f);
        wsm[((k / 8) * 4 + n / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2half(v);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    const uint32_t tw = tb + ((uint32_t)(warp * 32) << 16);
    const int row = threadIdx.x;
    uint32_t a[8];
    for (int c = 0; c < 8; ++c) {                          // A[row][k] = row/8 + k/16  (exact in fp16)
        __half2 h = __floats2half2_rn(row * 0.125f + (2 * c) * 0.0625f, row * 0.125f + (2 * c + 1) * 0.0625f);
        a[c] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st8(tw + 32, a);
    tc_wait_st(); tc_fence_before(); __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            tc_fence_after();
            const uint32_t idesc_f16acc = (0u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            mma_ts(tb, tb + 32, make_b_desc(smem_u32(wsm), 512, 128), idesc_f16acc, 0u);
            tc_commit(smem_u32(&bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    uint32_t raw[16], pk[8];
    tmem_ld16(tw, raw);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3]), "=r"(pk[4]), "=r"(pk[5]), "=r"(pk[6]), "=r"(pk[7]) : "r"(tw) : "memory");
    tc_wait_ld();
    for (int i = 0; i < 16; ++i) out[row * 24 + i] = raw[i];
    for (int i = 0; i < 8; ++i) out[row * 24 + 16 + i] = pk[i];
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u) : "memory"); }
}

static int run_f16acc() {
    uint32_t* d; CK(cudaMalloc(&d, 128 * 24 * 4)); CK(cudaMemset(d, 0xff, 128 * 24 * 4));
    f16acc_probe<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    uint32_t h[128 * 24];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    for (int row : {0, 1, 37, 100}) {
        printf("f16acc row %3d (A[row][k] = row/8 + k/16; D[n] = A[n] (n<16), 2A[n-16])\n  raw x16 cols 0..15:", row);
        for (int i = 0; i < 16; ++i) printf(" %08x", h[row * 24 + i]);
        printf("\n  pack::16b x8      :");
        for (int i = 0; i < 8; ++i) printf(" %08x", h[row * 24 + 16 + i]);
        printf("\n  as halves         :");
        for (int i = 0; i < 8; ++i) { __half2 v = *reinterpret_cast<__half2*>(&h[row * 24 + 16 + i]); printf(" (%g,%g)", __half2float(v.x), __half2float(v.y)); }
        printf("\n");
    }
    cudaFree(d);
    return 0;
}

int main() {
    run_f16acc();
    for (int grid : {148}) {
        for (int w : {1, 4, 8, 12, 16}) run("ld.x16", T_LD16, w, 0, grid, 4 * 16 * 32 * 4.0, "B");
        for (int w : {4, 8, 12}) run("ld.x32", T_LD32, w, 0, grid, 2 * 32 * 32 * 4.0, "B");
        for (int w : {1, 4, 8, 12}) run("st.x8", T_ST8, w, 0, grid, 8 * 8 * 32 * 4.0, "B");
        for (int w : {1, 4, 8, 12, 16}) run("tanh.f32", T_TANH32, w, 0, grid, 16 * 32.0, "tanh");
        for (int w : {1, 4, 8, 12, 16}) run("tanh.f16x2", T_TANH16, w, 0, grid, 16 * 64.0, "tanh");
        for (int w : {1, 4, 8, 12, 16}) run("fma.f32x2", T_FMA2, w, 0, grid, 16 * 32.0, "fma2-instr-lanes");
        for (int w : {1, 4, 8, 12, 16}) run("cvt.f16x2", T_PACK, w, 0, grid, 16 * 32.0, "cvt-instr-lanes");
        for (int w : {4, 8, 12, 16}) run("act-mix", T_MIX, w, 0, grid, 16 * 32.0, "activations");
        {
            const char* names[28] = {"tanh x8", "cvt x8", "fma2 x8", "hfma2 x8", "prmt x8", "tanh8+cvt8", "tanh8+fma2_8", "cvt8+fma2_8",
                                     "tanh8+hfma2_8", "cvt8+hfma2_8", "fma2_8+hfma2_8", "tanh8+cvt8+fma2_8", "tanh4+cvt6+fma2_8",
                                     "cvt8+prmt8", "tanh8+prmt8", "act-math f32x2 (16 act)", "act-math half2 (16 act)", "form0 f32x2", "form1 scalar f32", "form2 half2 tail", "form3 half2 tangent product", "form4 all half2", "ovl tanh8", "ovl ffma48", "ovl tanh8+ffma16", "ovl tanh8+ffma32", "ovl tanh8+ffma48", "ovl tanh8+ffma64"};
            for (int c = 22; c < 28; ++c) for (int w : {4, 12, 16}) run(names[c], T_TANH_PACK, w, c, grid, 1.0, "(per-thread inner iterations x warps)/clk");
        }
        for (int k : {1, 2, 3, 6, 8, 12}) run("mma", T_MMA, 1, k, grid, 0, "");
        for (int w : {4, 8, 12}) for (int k : {6, 8}) run("chain", T_CHAIN, w, k, grid, 1.0, "rounds");
        for (int w : {4, 8, 12}) for (int k : {6, 8}) run("chain2", T_CHAIN2, w, k, grid, 1.0, "rounds");
    }
    return 0;
}
