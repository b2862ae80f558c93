// Saturated rate of the SHIPPED activation pass in isolation: W warps per CTA (1..4 per scheduler) each loop over exactly the
// per-round worker code of flow_tc.cu -- tcgen05.ld of z (fp32) and du, dv (packed fp16), 32 x (tanh, f32x2 h, two packs,
// five half2 ops), tcgen05.st of the three operands, wait::st -- with no MMA, barrier or mbarrier in between.  What it
// measures: cycles per pass per scheduler when 1..4 warps compete, i.e. the throughput bound of the pass's own
// instruction mix (XU 256 cycles, FMA-heavy ~230, issue slots ~200 per pass).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pass_rate pass_rate.cu
#include <cstdint>
#include <cstdio>
typedef unsigned long long f32x2;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t pack_h2(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float tanh_approx(float x) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t hadd2_(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hsub2_(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmul2_(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hfma2_(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8_pack16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) { tmem_st4(taddr, r[0], r[1], r[2], r[3]); tmem_st4(taddr + 4, r[4], r[5], r[6], r[7]); }
// FORM 0: shipped.  1: h from two scalar FFMAs (FMA-lite capable) instead of one FFMA2.  2: (1 - t), (1 + t) from four scalar FADDs and two
// packs instead of pack(t) + two HADD2.  3: both.
#ifndef FORM
#define FORM 0
#endif
__device__ __forceinline__ void activate16(const float* z, const uint32_t* du, const uint32_t* dv, uint32_t* ph, uint32_t* pu, uint32_t* pv) {
    constexpr uint32_t kOneH2 = 0x3C003C00u;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const float t0 = tanh_approx(z[j]), t1 = tanh_approx(z[j + 1]);
        uint32_t h16;
        if (FORM & 1) h16 = pack_h2(fmaf(z[j], t0, z[j]), fmaf(z[j + 1], t1, z[j + 1]));
        else { const f32x2 zh = pk2(z[j], z[j + 1]); h16 = pack_h2(fma2(zh, pk2(t0, t1), zh)); }
        uint32_t a16, b16;
        if (FORM & 2) { a16 = pack_h2(1.0f - t0, 1.0f - t1); b16 = pack_h2(1.0f + t0, 1.0f + t1); }
        else { const uint32_t t16 = pack_h2(t0, t1); a16 = hsub2_(kOneH2, t16); b16 = hadd2_(kOneH2, t16); }
        const uint32_t s16 = hfma2_(h16, a16, b16);
        ph[j >> 1] = h16; pu[j >> 1] = hmul2_(s16, du[j >> 1]); pv[j >> 1] = hmul2_(s16, dv[j >> 1]);
    }
}
__global__ void __launch_bounds__(512, 1) k(int nwarps, int with_tmem, long long* out) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t tg = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    if (warp < 16) { for (int c = 0; c < 128; c += 4) tmem_st4(tg + c, 0, 0, 0, 0); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    __syncthreads();
    if (warp < nwarps) {
        float za[16], zb[16];
        uint32_t ua[8], va[8], ub[8], vb[8], ph[8], pu[8], pv[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) { za[i] = 0.01f * (lane + i); zb[i] = -0.02f * (lane + i); }
#pragma unroll
        for (int i = 0; i < 8; ++i) { ua[i] = va[i] = 0x38003800u + i; ub[i] = vb[i] = 0x34003400u + i; }
        uint32_t sinkv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 2000; ++it) {
            const bool LD = with_tmem & 1, ST = with_tmem & 2;
            if (LD) {
                tmem_ld16(tg, za); tmem_ld8_pack16(tg + 32, ua); tmem_ld8_pack16(tg + 64, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld16(tg + 16, zb); tmem_ld8_pack16(tg + 48, ub); tmem_ld8_pack16(tg + 80, vb);
            } else {                                   // inputs change through one real instruction each (xor with the previous outputs' parity bit)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    za[i] = __uint_as_float(__float_as_uint(za[i]) ^ (sinkv[i & 7] & 1u));
                    zb[i] = __uint_as_float(__float_as_uint(zb[i]) ^ (sinkv[(i + 3) & 7] & 1u));
                }
            }
            activate16(za, ua, va, ph, pu, pv);
            if (ST) { tmem_st8(tg + 96, ph); tmem_st8(tg + 64, pu); tmem_st8(tg + 112, pv); }
            else {                                     // a real (cheap) consumer: 8 three-input xors per half
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sinkv[i]) : "r"(ph[i]), "r"(pu[i]), "r"(pv[i]));
            }
            if (LD) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            activate16(zb, ub, vb, ph, pu, pv);
            if (ST) { tmem_st8(tg + 104, ph); tmem_st8(tg + 72, pu); tmem_st8(tg + 120, pv); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sinkv[i]) : "r"(ph[i]), "r"(pu[i]), "r"(pv[i]));
            }
        }
        const long long t1 = clock64();
        if (lane == 0) out[warp] = (t1 - t0) / 2000;
        if (za[0] + zb[0] == 12345.f) out[31] = ph[0] + sinkv[0] + sinkv[1] + sinkv[2] + sinkv[3] + sinkv[4] + sinkv[5] + sinkv[6] + sinkv[7];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
int main() {
    long long* out; cudaMallocManaged(&out, 32 * sizeof(long long));
    const char* names[4] = {"math only (registers)     ", "tcgen05.ld + math         ", "math + tcgen05.st         ", "full pass (ld, math, st)  "};
    printf("FORM %d\n", FORM);
    for (int with_tmem = 3; with_tmem >= 3; --with_tmem)
        for (int w : {4, 8, 12, 16}) {
            k<<<148, 512>>>(w, with_tmem, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long mx = 0; for (int i = 0; i < w; ++i) mx = out[i] > mx ? out[i] : mx;
            printf("%s  %2d warps (%d per scheduler): %5lld cycles per pass per warp -> %6.1f cycles per pass per scheduler\n",
                   names[with_tmem], w, w / 4, mx, (double)mx / (w / 4));
        }
    return 0;
}
