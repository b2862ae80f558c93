// PREC_TC16 path: the per-query flow on 5th-generation tensor cores (tcgen05) for sm_100a.
//
// One persistent CTA per SM, kGroups worker groups of 128 threads (4 warps; thread <-> TMEM lane <->
// query row).  Each group owns one 128-query tile at a time and runs the WHOLE T-step flow for it on
// chip; the groups are independent pipelines that overlap each other's tensor-core round trips
// (TMEM holds exactly four tiles: 4 x 128 of the 512 columns).
//
//   per tile:  PE5(wi) -> fp16 -> TMEM A1[:,8:30];  base net, x0 (Philox or replay), p0   (fp32)
//   per step:  state (hi/lo fp16 split) -> A1[:,0:8];  tangent seeds d(state)/dx -> A_u, A_v (K chunk 0)
//              round 0      :  D_z = A1 . W1^T (K=32),  D_u = A_u . W1^T, D_v = A_v . W1^T (K=16)
//              rounds 1..H-1:  D_z, D_u, D_v = A_{h,u,v} . Wl^T                              (K=32, N=32)
//              after every round the SAME code runs:  tcgen05.ld z, du, dv ; t = tanh(z/2) ;
//                   h = silu(z), s = silu'(z) ; A_h = h, A_u = s du, A_v = s dv  (fp16, tcgen05.st)
//              output round :  D = A . Wout^T (N=16) ; d, dd/dx0, dd/dx1 (6 floats) ;
//                   det J, R, Euler update in fp32 registers
//   epilogue:  base log-prob (pdf mode), domain mapping + Jacobian, store wo / pdf
//
// The first layer's tangents are produced by the tensor core as well: the seed operand is the derivative
// of the first-layer input row w.r.t. the state (a unit vector for x, theta; (cos phi, -sin phi) on the
// sin/cos columns for phi), so D_u = W1 . d(in)/dx0 and every round, first or hidden, is the same
// instruction stream -- there is no per-neuron table lookup and the activation loop exists once.
//
// Synchronisation per round (measured chain ~320 cycles beyond the loads/stores, profiles/microbench):
//   tcgen05.wait::st ; tcgen05.fence::before_thread_sync ; bar.sync (group, 128 threads) ;
//   one elected lane of the group's first warp: fence::after, 6-8 x tcgen05.mma, tcgen05.commit -> mbarrier ;
//   all 128 threads: mbarrier.try_wait (hardware-suspended) ; fence::after ; tcgen05.ld
// All MMA operands are warp-uniform (TMEM base through a shuffle, compile-time offsets), so the MMAs issue
// back to back from uniform registers.
//
// Operands: A (activations; value + two tangent columns = three M=128 row blocks sharing B) lives in
// TMEM as fp16 (tcgen05.mma ".ts" form), written by the worker threads with tcgen05.st -- activations
// never touch shared memory or HBM.  B (weights) is resident in shared memory for the CTA's lifetime:
// the packed blob's fp16 image is ALREADY the UMMA canonical K-major layout, staged with one
// cp.async.bulk (TMA).  Every weight matrix is stored as fp16 hi + fp16 lo; the value path multiplies
// by both (weights effectively ~22 bits: weight rounding was the dominant fp16 error), the tangent path
// by hi only.  Value accumulators (z) and the output round are fp32 in TMEM; the hidden-layer TANGENT
// accumulators are fp16 (idesc D format f16) and are read back two per register with tcgen05.ld .pack::16b,
// so s2 and the tangent products run as half2 instructions and need no conversion.  x, det, R, pdf stay
// fp32 in registers for all T steps.
//
// tanh-form sigmoid: hidden-layer weights are pre-scaled by 1/2 (exact), so the MMA yields zh = z/2 and
// duh = du/2:   t = tanh(zh);  silu(z) = zh + zh t;  2 silu'(z) = (1 + t) + silu(z) (1 - t);
// u_out = 2 silu'(z) * duh  -> one MUFU op per activation and no rescaling.  Per neuron PAIR the math is
// 2 MUFU.TANH + 1 fma.f32x2 (h) + 2 cvt.f16x2 (h, t) + 5 half2 ops (1-t, 1+t, s2, s2 du, s2 dv): measured
// MUFU-bound in isolation (16 tanh / clk / SM; profiles/microbench "form2"), where the all-fp32 formulation
// was dispatch-bound at 10.7.
//
// TMEM map (512 columns allocated; per group 128 columns at g*128):
//   [  0, 32)  D_z         fp32 value accumulators (output round: 16 columns)
//   [ 32, 64)  D_u         tangent accumulators: fp16 in the hidden rounds, fp32 (16 columns) in the output round
//   [ 64, 96)  D_v         same; its first 16 columns are ALSO A_u (see below)
//   [ 64, 80)  A_u         fp16 operand, K=32 -> 16 columns
//   [ 96,112)  A_h / A1    fp16 operand of the hidden rounds; at every step start it is rewritten as the first-layer
//                          operand A1 (cols 0..3 state, 4..15 PE5(wi) re-read from the tile's shared-memory record)
//   [112,128)  A_v         fp16 operand
// (BSDFDIFF_TC_GROUPS=5 builds a 96-column map with five tiles in flight -- v-tangent operand through shared memory --
// which is bit-identical and no faster: DESIGN.md 4.1.)
// A_u and D_v overlap so that a tile needs 128 instead of 144 columns (four tiles in flight, not three).  A round
// issues its MMAs in the order  u (reads A_u) ; z (four MMAs) ; v (writes D_v):  tcgen05.mma instructions of one
// thread execute in issue order, so A_u has been consumed five MMAs before the first write to D_v; threads only
// write A_u after their tcgen05.ld of D_v has completed (tcgen05.wait::ld).  BSDFDIFF_TC_NOALIAS builds the
// non-overlapping 144-column map (three groups) that the aliased build is checked against bit for bit.
#include "common.cuh"

#include <type_traits>

namespace bsdfdiff {

#ifndef BSDFDIFF_TC_GROUPS
#ifdef BSDFDIFF_TC_NOALIAS
#define BSDFDIFF_TC_GROUPS 3      // the 144-column map: only three tiles fit
#else
#define BSDFDIFF_TC_GROUPS 4      // tuning builds only: 1 or 2 groups isolate the per-round latency chain
#endif
#endif
constexpr int kGroups = BSDFDIFF_TC_GROUPS;          // tiles in flight per CTA ("pipes"): one TMEM column block, one d_ready barrier each
// BSDFDIFF_TC_DUO: every worker thread owns the same row of TWO tiles and alternates between them, so a tile's
// tensor-core round trip runs under the OTHER tile's activation pass of the same thread (two worker groups x two
// tiles instead of four groups x one tile): no warp ever sits in bar.sync or on an mbarrier with nothing to do, and
// 12 warps at up to 168 registers replace 20 at 96.  0 = the round-1 structure (one tile per thread), kept for A/B builds.
#ifndef BSDFDIFF_TC_DUO
#define BSDFDIFF_TC_DUO 0        // measured (profiles/r2b_variants.txt): 4.5-5.5 ms against 3.05 ms for 16.7 M disk queries -- two
#endif                           // worker warps per scheduler cannot cover the activation math's own latencies; tuning build only
constexpr bool kDuo = BSDFDIFF_TC_DUO != 0;
static_assert(!kDuo || kGroups == 4, "duo workers: two groups x two tiles");
constexpr int kWGroups = kDuo ? kGroups / 2 : kGroups;   // worker thread groups (128 threads each)
// who issues a tile's MMAs once its 128 rows are stored (duo only): 2 = the warp that arrives LAST (shared-memory
// arrival counter; nobody waits for anybody), 1 = a rotating warp waits in bar.sync while the others bar.arrive and
// move on, 0 = all four warps bar.sync (round-1 behaviour)
#ifndef BSDFDIFF_TC_ISSUE
#define BSDFDIFF_TC_ISSUE 2
#endif
#ifndef BSDFDIFF_TC_SKEW
#define BSDFDIFF_TC_SKEW 2       // turns by which a thread's second tile trails its first (so their short output passes do not meet)
#endif
// BSDFDIFF_TC_ISSUER (tuning build): one dedicated MMA-issuer warp per group.  Worker warps never block in a barrier:
// they bar.arrive in the middle of an activation pass (first 16 neurons stored and every accumulator read) and at its
// end; the group's issuer warp bar.syncs on the same named barriers and issues K-chunk 0 of the next round's u and z
// MMAs under the workers' second-half math, the rest (+ commit) at the end.  24 warps -> 80 registers per thread.
#ifndef BSDFDIFF_TC_ISSUER
#define BSDFDIFF_TC_ISSUER 0
#endif
constexpr bool kIssuer = BSDFDIFF_TC_ISSUER != 0;
static_assert(!kIssuer || (!kDuo && BSDFDIFF_TC_GROUPS == 4), "issuer warps: the four-group, one-tile-per-thread structure");
constexpr int kWorkerThreads = kWGroups * 128;
constexpr int kProducerWarps = 4;                  // one 128-thread producer group: thread <-> query row of a tile
constexpr int kIssuerWarps = kIssuer ? kGroups : 0;
constexpr int kTcThreads = kWorkerThreads + 32 * kProducerWarps + 32 * kIssuerWarps;
#ifndef BSDFDIFF_TC_SLOTS
#define BSDFDIFF_TC_SLOTS (BSDFDIFF_TC_GROUPS + 2)
#endif
#ifndef BSDFDIFF_TC_BACKOFF_NS
#define BSDFDIFF_TC_BACKOFF_NS 2000
#endif
constexpr int kSlots = BSDFDIFF_TC_SLOTS;                // prologue ring: every group holds its slot for the whole tile (PE5 is
                                                   // re-read each step), the producer runs up to two tiles ahead
// per-query record handed from the producer to the worker thread of the same row (field-major in shared memory,
// slot[field][row], so both sides access consecutive words)
constexpr int kFPe = 0;                            // 11 words: PE5(wi) as packed fp16 pairs (A1 columns 4..14)
constexpr int kFX = 11;                            // x0, x1: start state (base sample, replayed noise, or wo)
constexpr int kFP0 = 13;                           // base density at the start state (sample mode)
constexpr int kFBp = 14;                           // 4 base-net outputs (pdf mode evaluates the base density at the end)
constexpr int kFWiz = 18;                          // wi_z, wo (pdf-mode masks)
constexpr int kFWo = 19;
constexpr int kFIdx = 22;                          // multi-material launches: wavefront row of this tile row (int bits), -1 = padding row
constexpr int kFields = 23;
constexpr int kTile = 128;
#if BSDFDIFF_TC_GROUPS == 5
// Five tiles in flight: 96 columns per tile.  The v-tangent operand goes through shared memory (SS-form MMA), A_h sits
// inside D_u and A_u inside D_v; a round issues z (reads A_h), u (reads A_u, writes D_u over A_h), v (reads smem,
// writes D_v over A_u) -- tcgen05.mma of one thread execute in issue order.  64-wide forward rounds: D_z 64 + A_h 32.
constexpr bool kSmemAv = true;
constexpr int kColsPerGroup = 96;
constexpr int kColDz = 0, kColDu = 32, kColDv = 64, kColAu = 64, kColAh32 = 32, kColAh64 = 64, kColAv = 0;
#elif defined(BSDFDIFF_TC_NOALIAS)
constexpr bool kSmemAv = false;
constexpr int kColsPerGroup = 144;
constexpr int kColDz = 0, kColDu = 32, kColDv = 64, kColAu = 96, kColAh32 = 112, kColAh64 = 112, kColAv = 128;
#else
constexpr bool kSmemAv = false;
constexpr int kColsPerGroup = 128;
constexpr int kColDz = 0, kColDu = 32, kColDv = 64, kColAu = 64, kColAh32 = 96, kColAh64 = 96, kColAv = 112;
#endif
template <int H> __host__ __device__ constexpr int col_ah() { return H == 64 ? kColAh64 : kColAh32; }
constexpr int kAvBytes = kTile * 64;                // one tile's v-tangent operand in shared memory: 128 rows x K = 32 halves
constexpr int kTmemCols = 512;
static_assert(kGroups * kColsPerGroup <= kTmemCols, "TMEM budget");
static_assert(kSlots > kGroups && kSlots < 2 * kGroups + 1, "slot ring: one wrap per worker iteration at most");

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

// BSDFDIFF_TC_TRACE (tuning builds): lane 0 of every warp of CTA 0 stamps %clock64 at fixed points of its rounds
// kTraceFirst .. kTraceFirst + kTraceRounds - 1 (steady state); read back with bsdfdiff_debug_trace and analysed by
// profiles/trace_timeline.py.  Events: 0 woke from the MMA wait, 1 first half loaded, 2 first half computed + stored,
// 3 second half computed + stored, 4 stores landed (wait::st), 5 left the group barrier, 6 MMAs issued (issuer warp)
constexpr int kTraceRounds = 96, kTraceFirst = 160, kTraceEvents = 8, kTraceWarps = 24;
#ifdef BSDFDIFF_TC_TRACE
__device__ unsigned long long g_tc_trace[kTraceWarps][kTraceRounds][kTraceEvents];
#define TC_TRACE(ev)                                                                                          \
    do {                                                                                                      \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_r >= kTraceFirst && trace_r < kTraceFirst + kTraceRounds) \
            g_tc_trace[threadIdx.x >> 5][trace_r - kTraceFirst][ev] = clock64();                              \
    } while (0)
#else
#define TC_TRACE(ev) do { } while (0)
#endif

// Wait for the phase with the given parity.  try_wait suspends the warp in hardware instead of spinning, but it
// also wakes on unrelated mbarrier traffic of the CTA; BACKOFF_NS > 0 (the producer, which runs a whole tile ahead)
// adds a nanosleep between probes so a long wait does not burn issue slots of the computing warps.  The loop is
// bounded (0x400000 probes of <= 20 us): a protocol bug records the fault and aborts the launch instead of hanging
// the GPU.  The fault path (flag + trap) sits inside the asm, so the hot path carries no status flag and no
// reconvergence scaffolding (-2 % run time against the version that returned a flag).
template <int BACKOFF_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (BACKOFF_NS > 0) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
            "mov.u32 n, 0x400000;\n"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
            "vote.sync.all.pred q, p, 0xffffffff;\n\t"
            "@q bra.uni DONE_%=;\n\t"
            "nanosleep.u32 %5;\n\t"
            "sub.u32 n, n, 1;\n\t"
            "setp.ne.u32 p, n, 0;\n\t"
            "@p bra.uni WAIT_%=;\n\t"
            "st.global.u32 [%3], %4;\n\t"
            "trap;\n"
            "DONE_%=:\n\t}"
            :: "r"(bar), "r"(parity), "r"(20000u), "l"(&g_tc_timeout_flag), "r"(1u), "n"(BACKOFF_NS) : "memory");
    } else {
        // (the exit branch goes through a warp vote: ptxas then knows the loop is warp-uniform and keeps the TMEM / barrier
        // addresses of the surrounding code in uniform registers instead of re-deriving them with R2UR at every use)
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
            "mov.u32 n, 0x400000;\n"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
            "vote.sync.all.pred q, p, 0xffffffff;\n\t"
            "@q bra.uni DONE_%=;\n\t"
            "sub.u32 n, n, 1;\n\t"
            "setp.ne.u32 p, n, 0;\n\t"
            "@p bra.uni WAIT_%=;\n\t"
            "st.global.u32 [%3], %4;\n\t"
            "trap;\n"
            "DONE_%=:\n\t}"
            :: "r"(bar), "r"(parity), "r"(20000u), "l"(&g_tc_timeout_flag), "r"(1u) : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::f16 (fp16 operands, fp32 accumulate); ACC is a compile-time flag
template <int ACC>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T (A through a shared-memory descriptor)
template <int ACC>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC) : "memory");
}
// eight operand words (16 halves = two 8-half core-matrix rows) of this thread's row into the shared-memory A image:
// K-major canonical layout, core column c at +c * 2048 bytes (16 row groups x 128 B)
__device__ __forceinline__ void smem_st_a16(uint32_t row_addr, int chunk, const uint32_t* r) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row_addr + (uint32_t)chunk * 2048u), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row_addr + (uint32_t)(chunk + 1) * 2048u), "r"(r[4]),
                 "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// 16 columns of fp16 accumulators (one half per 32-bit column) -> 8 registers, two neurons per register
__device__ __forceinline__ void tmem_ld8_pack16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4_pack16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.pack::16b.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t x, y;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(taddr) : "memory");
    a = __uint_as_float(x); b = __uint_as_float(y);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Eight operand words of this thread's row.  Issued as two .x4 stores: measured on B200, an .x8 (or .x16) store
// holds the issuing warp ~20-45 cycles, an .x4 store ~5 (profiles/microbench/mma_issue: 6 x .x8 = 130-300 cycles,
// 12 x .x4 = 56 cycles for the same bytes).
#ifndef BSDFDIFF_TC_ST_X8
#define BSDFDIFF_TC_ST_X8 0
#endif
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    if (BSDFDIFF_TC_ST_X8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    } else {
        tmem_st4(taddr, r[0], r[1], r[2], r[3]);
        tmem_st4(taddr + 4, r[4], r[5], r[6], r[7]);
    }
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// fp32 value of the low / high half of a packed fp16 pair
__device__ __forceinline__ float h2_lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xffffu))); }
__device__ __forceinline__ float h2_hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }

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

// packed fp16 pairs
__device__ __forceinline__ uint32_t hadd2_(uint32_t a, uint32_t b) {
    uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t hsub2_(uint32_t a, uint32_t b) {
    uint32_t d; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t hmul2_(uint32_t a, uint32_t b) {
    uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t hfma2_(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32 or fp16, A/B fp16, both K-major, M=128 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int N, bool d_f32) {
    return (d_f32 ? (1u << 4) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// One activation pass over 16 neurons of this thread's row.  Inputs: zh = z/2 (fp32, from TMEM) and the packed
// fp16 tangent pre-activations duh = du/2, dvh = dv/2.  Outputs are the next round's fp16 operand words:
//   h = silu(z) = zh + zh t,   s2 = 2 silu'(z) = (1 + t) + h (1 - t),   u = s2 duh,   v = s2 dvh,   t = tanh(zh)
//   ACT 1: tanh.approx.f32; h in fp32 (f32x2 fma), s2 and the products in half2
//   ACT 0: fp32 exp form for h and s2 (cross-check variant), products in half2
// ------------------------------------------------------------------------------------------------
template <bool TANGENTS, int ACT>
__device__ __forceinline__ void activate16(const float* z, const uint32_t* du, const uint32_t* dv,
                                           uint32_t* ph, uint32_t* pu, uint32_t* pv) {
    constexpr uint32_t kOneH2 = 0x3C003C00u;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        uint32_t h16, s16 = 0u;
        if (ACT == 0) {
            float hh[2], ss[2];
            const float zz[2] = {2.0f * z[j], 2.0f * z[j + 1]};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float s = __fdividef(1.0f, 1.0f + __expf(-zz[i]));
                hh[i] = zz[i] * s;
                ss[i] = 2.0f * s * fmaf(zz[i], 1.0f - s, 1.0f);
            }
            h16 = pack_h2(hh[0], hh[1]);
            if (TANGENTS) s16 = pack_h2(ss[0], ss[1]);
        } else {
            const float t0 = tanh_approx(z[j]), t1 = tanh_approx(z[j + 1]);
            const f32x2 zh = pk2(z[j], z[j + 1]);
            h16 = pack_h2(fma2(zh, pk2(t0, t1), zh));                 // zh (1 + t)
            if (TANGENTS) {
                const uint32_t t16 = pack_h2(t0, t1);
                s16 = hfma2_(h16, hsub2_(kOneH2, t16), hadd2_(kOneH2, t16));   // (1 + t) + h (1 - t)
            }
        }
        ph[j >> 1] = h16;
        if (TANGENTS) {
            pu[j >> 1] = hmul2_(s16, du[j >> 1]);
            pv[j >> 1] = hmul2_(s16, dv[j >> 1]);
        }
    }
}

// BSDFDIFF_TC_PIPE4 (tuning build): the pass over 32 neurons as four chunks of 8, software-pipelined so that the MUFU ops of
// chunk c+1 execute under the half2 tail of chunk c (the shipped pass runs two bursts of 16 MUFU ops, each followed by
// the tail that depends on it).
#ifndef BSDFDIFF_TC_PIPE4
#define BSDFDIFF_TC_PIPE4 0
#endif
__device__ __forceinline__ void tanh8(const float* z, float* t) {
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = tanh_approx(z[j]);
}
__device__ __forceinline__ void tail8(const float* z, const float* t, const uint32_t* du, const uint32_t* dv,
                                      uint32_t* ph, uint32_t* pu, uint32_t* pv) {
    constexpr uint32_t kOneH2 = 0x3C003C00u;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const f32x2 zh = pk2(z[j], z[j + 1]);
        const uint32_t h16 = pack_h2(fma2(zh, pk2(t[j], t[j + 1]), zh));
        const uint32_t t16 = pack_h2(t[j], t[j + 1]);
        const uint32_t s16 = hfma2_(h16, hsub2_(kOneH2, t16), hadd2_(kOneH2, t16));
        ph[j >> 1] = h16;
        pu[j >> 1] = hmul2_(s16, du[j >> 1]);
        pv[j >> 1] = hmul2_(s16, dv[j >> 1]);
    }
}

// PE5(v): two MUFU sin/cos pairs (arguments are bounded: |v| <= pi for every domain) and four double-angle
// steps; abs error ~1e-5 at the highest frequency, well below the fp16 ulp of the operand it feeds.
__device__ __forceinline__ void pe5_fast(float v0, float v1, float* e) {
    float s0, c0, s1, c1;
    __sincosf(v0, &s0, &c0);
    __sincosf(v1, &s1, &c1);
    e[0] = v0; e[1] = v1;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        e[2 + 4 * k + 0] = s0; e[2 + 4 * k + 1] = s1; e[2 + 4 * k + 2] = c0; e[2 + 4 * k + 3] = c1;
        const float ns0 = (s0 + s0) * c0, ns1 = (s1 + s1) * c1;
        c0 = fmaf(c0, c0, -s0 * s0); c1 = fmaf(c1, c1, -s1 * s1);
        s0 = ns0; s1 = ns1;
    }
}

// ---- base distribution with MUFU-grade intrinsics (the tensor-core path tolerates ~1e-6 relative here; the
// fp32 parity path keeps the precise versions in common.cuh).  Same formulas as base_draw / base_logprob.
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.0f ? x : __logf(1.0f + __expf(x)); }
__device__ __forceinline__ float log_i0_fast(float x) {
    if (x < 3.75f) {
        float y = x * (1.0f / 3.75f); y = y * y;
        float r = 0.45813e-2f;
        r = fmaf(y, r, 0.360768e-1f); r = fmaf(y, r, 0.2659732f); r = fmaf(y, r, 1.2067492f);
        r = fmaf(y, r, 3.0899424f);   r = fmaf(y, r, 3.5156229f); r = fmaf(y, r, 1.0f);
        return __logf(r);
    }
    float y = __fdividef(3.75f, x);
    float r = 0.392377e-2f;
    r = fmaf(y, r, -0.1647633e-1f); r = fmaf(y, r, 0.2635537e-1f); r = fmaf(y, r, -0.2057706e-1f);
    r = fmaf(y, r, 0.916281e-2f);   r = fmaf(y, r, -0.157565e-2f); r = fmaf(y, r, 0.225319e-2f);
    r = fmaf(y, r, 0.1328592e-1f);  r = fmaf(y, r, 0.39894228f);
    return x - 0.5f * __logf(x) + __logf(r);
}
// (explicit __f*_rn intrinsics: the compiler may not contract or re-associate, so every build of this file -- the TMEM
// map variants in particular -- evaluates the density with the same roundings)
template <int DOMAIN>
__device__ __forceinline__ float base_logprob_fast(const float p[4], float x0, float x1) {
    if (DOMAIN == kDisk) {
        const float e0 = __fmul_rn(__fsub_rn(x0, p[0]), __expf(-p[2])), e1 = __fmul_rn(__fsub_rn(x1, p[1]), __expf(-p[3]));
        const float q = __fmaf_rn(e0, e0, __fmul_rn(e1, e1));
        return __fmaf_rn(-0.5f, q, __fsub_rn(-kLog2Pi, __fadd_rn(p[2], p[3])));
    }
    const float kappa = __fadd_rn(softplus_fast(p[3]), 1e-3f);
    const float e = __fdividef(__fsub_rn(x0, p[0]), __fadd_rn(__expf(p[1]), 1e-3f));
    const float loggau = __fmaf_rn(-0.5f, __fmul_rn(e, e), __fsub_rn(-0.5f * kLog2Pi, p[1]));
    const float logvon = __fsub_rn(__fmaf_rn(kappa, __cosf(__fsub_rn(x1, p[2])), -kLog2Pi), log_i0_fast(kappa));
    return __fadd_rn(loggau, logvon);
}
template <int DOMAIN>
__device__ __forceinline__ void base_draw_fast(const float p[4], unsigned long long seed, unsigned long long offset,
                                               long long index, float& x0, float& x1, const float* u = nullptr) {
    float ua, ub;
    if (u) {                                       // renderer-supplied uniforms (common.cuh: base_draw)
        ua = clamp_u01(u[0]); ub = clamp_u01(u[1]);
        noise_key(u, seed, offset);
        index = 0;
    } else {
        const uint4 r4 = philox_draw(seed, offset, index, 0);
        ua = u01(r4.x); ub = u01(r4.y);
    }
    float n0, n1;
    {
        const float r = sqrtf(-2.0f * __logf(ua));
        float s, c;
        __sincosf(6.283185307179586f * ub - 3.141592653589793f, &s, &c);   // argument in (-pi, pi)
        n0 = -r * c; n1 = -r * s;
    }
    if (DOMAIN == kDisk) {
        x0 = fmaf(n0, __expf(p[2]), p[0]);
        x1 = fmaf(n1, __expf(p[3]), p[1]);
        return;
    }
    x0 = fmaf(n0, __expf(p[1]) + 1e-3f, p[0]);
    const float kappa = softplus_fast(p[3]) + 1e-3f, mu = p[2];
    const float s = sqrtf(fmaf(4.0f * kappa, kappa, 1.0f));
    const float tau = 1.0f + s;
    const float rho = __fdividef(tau * 2.0f * kappa, (s + 1.0f) * (tau + sqrtf(2.0f * tau)));
    const float r = __fdividef(1.0f + rho * rho, 2.0f * rho);
    float phi = 0.0f;
    for (uint32_t round = 1; round <= 64; ++round) {                // acceptance >= ~0.66 per round
        const uint4 q = philox_draw(seed, offset, index, round);
        const float z = __cosf(3.141592653589793f * u01(q.x));
        const float f = __fdividef(1.0f + r * z, r + z);
        const float cc = kappa * (r - f);
        const float u2 = u01(q.y);
        phi = ((q.z & 0x80000000u) ? 1.0f : -1.0f) * acosf(fminf(fmaxf(f, -1.0f), 1.0f));
        if ((cc * (2.0f - cc) - u2 > 0.0f) || (__logf(__fdividef(cc, u2)) + 1.0f - cc >= 0.0f)) break;
    }
    float y = phi + 3.14159265358979f + mu;
    y = y - 6.28318530717959f * floorf(y * 0.159154943091895f);   // python-style modulo
    x1 = y - 3.14159265358979f;
}

// Append row i to the fix-up list (warp-aggregated: one atomic per warp that has any flagged row).  Called by all 32
// lanes of a converged worker warp.
// Multi-material launches keep one list per material (a warp's rows share a tile, hence a material): material m's
// list lives at fix_list[seg_off[m] ...] -- it can never outgrow the material's own row count -- with its length in
// fix_count[m], so the fix-up pass can stage one weight set per chunk of rows.
__device__ __forceinline__ void flag_for_fixup(const FlowParams& P, bool flag, long long i, int mat = -1,
                                               float x0_init = 0.0f, float x1_init = 0.0f) {
    const uint32_t mask = __ballot_sync(0xffffffffu, flag);
    if (mask == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned int* count = (mat >= 0) ? P.fix_count + mat : P.fix_count;
    const unsigned int seg = (mat >= 0) ? P.seg_off[mat] : 0u;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned int)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) {
        const unsigned int at = seg + base + __popc(mask & ((1u << lane) - 1u));
        P.fix_list[at] = (unsigned int)i;
        if (P.fix_x0) reinterpret_cast<float2*>(P.fix_x0)[at] = make_float2(x0_init, x1_init);   // the base sample to replay
    }
}

struct TcSmem {
    unsigned long long d_ready[kGroups];
    unsigned int arrive[kGroups];             // duo, issue mode 2: warps that have stored their rows of the pipe's current round
    unsigned long long w_bar[kGroups];        // weights landed (single material: [0] only; multi: one image per group)
    unsigned long long full[kSlots];          // producer -> worker group: the slot's records are written
    unsigned long long empty[kSlots];         // worker group -> producer: the slot has been read
    uint32_t tmem_base;
    uint32_t pad[3];
    // base net re-laid for 128-bit broadcast loads: w1t[k][j] (k < 14, j < 16), b1[16], wot[j][c] (c < 4), bo[4]
    __align__(16) float bw1t[kPE3 * 16];
    __align__(16) float bb1[16];
    __align__(16) float bwot[16 * 4];
    __align__(16) float bbo[4];
    __align__(16) float slot[kSlots][kFields][kTile];
    __align__(128) unsigned char av[kSmemAv ? kGroups : 1][kSmemAv ? kAvBytes : 128];   // v-tangent operands (5-tile map)
    __align__(128) unsigned char w16[128];    // hi+lo operand images of every layer: dynamic tail, hdr->f16_bytes long
};
static inline size_t tc_image_bytes(int H, int n_hidden) {
    return (sizeof(__half) * (size_t)f16_image_halves(H, n_hidden) + 127) / 128 * 128;
}
static inline size_t tc_smem_bytes(int H, int n_hidden, bool multi) {
    return offsetof(TcSmem, w16) + (multi ? kGroups : 1) * tc_image_bytes(H, n_hidden);
}

// base net p = Wo silu(W1 PE3(e) + b1) + bo from the shared-memory copy (rendering/utils/model.py:382-386)
__device__ __forceinline__ void base_eval_smem(const TcSmem& S, const float* e, float p[4]) {
    f32x2 z[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = reinterpret_cast<const float4*>(S.bb1)[q];
        z[2 * q] = pk2(b.x, b.y); z[2 * q + 1] = pk2(b.z, b.w);
    }
#pragma unroll
    for (int k = 0; k < kPE3; ++k) {
        const f32x2 ek = pk2(e[k], e[k]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = reinterpret_cast<const float4*>(S.bw1t)[k * 4 + q];
            z[2 * q] = fma2(ek, pk2(w.x, w.y), z[2 * q]);
            z[2 * q + 1] = fma2(ek, pk2(w.z, w.w), z[2 * q + 1]);
        }
    }
    const float4 bo = *reinterpret_cast<const float4*>(S.bbo);
    f32x2 p01 = pk2(bo.x, bo.y), p23 = pk2(bo.z, bo.w);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        float z0, z1;
        upk2(z[j >> 1], z0, z1);
        // silu(z) = z/2 + z/2 tanh(z/2): one MUFU op per neuron (the exp form needs ex2 + rcp)
        const float zh0 = 0.5f * z0, zh1 = 0.5f * z1;
        const float h0 = fmaf(zh0, tanh_approx(zh0), zh0);
        const float h1 = fmaf(zh1, tanh_approx(zh1), zh1);
        const float4 w0 = reinterpret_cast<const float4*>(S.bwot)[j], w1 = reinterpret_cast<const float4*>(S.bwot)[j + 1];
        const f32x2 hh0 = pk2(h0, h0), hh1 = pk2(h1, h1);
        p01 = fma2(hh0, pk2(w0.x, w0.y), p01); p23 = fma2(hh0, pk2(w0.z, w0.w), p23);
        p01 = fma2(hh1, pk2(w1.x, w1.y), p01); p23 = fma2(hh1, pk2(w1.z, w1.w), p23);
    }
    upk2(p01, p[0], p[1]); upk2(p23, p[2], p[3]);
}

// Issue one round of MMAs for a group (executed by ONE thread; every operand is warp-uniform).  Each layer's
// operand image = HI [N x 32] then LO [N x 32] (W = hi + lo): the value path accumulates A.hi^T + A.lo^T,
// tangents use hi only.  type 0: first layer (tangent seeds are K=16 operands in A_u/A_v chunk 0),
// 1..NH-1: hidden layer, NH: output layer (N = 16, fp32 accumulators throughout).
// Order: u, z, v -- A_u is read first and D_v (which overlaps A_u) is written last.
// Split issue of one round (issuer-warp builds; 32-wide tangent rounds, aliased TMEM map): PART 1 = K-chunk 0 of the u and z
// products (their operands are complete once the first 16 neurons of the previous pass are stored; D_u and D_z have been
// read by then), PART 2 = everything else + commit.  Order inside PART 2: u (chunk 1) before the v products, which write
// D_v over A_u.  The first-layer round (type 0) has no early part: its operand is written in one go.
template <int PART>
__device__ __forceinline__ void issue_round_split(int type, int NH, uint32_t tg, uint64_t b_first, uint64_t b_out, uint32_t bar) {
    constexpr int H = 32;
    constexpr uint32_t idz = make_idesc(H, true), idt = make_idesc(H, false), ido = make_idesc(16, true);
    constexpr uint32_t kChunk = (2u * H * 16u) >> 4, kLoHid = (H * H * 2u) >> 4;
    constexpr uint32_t kFirstBytes = 2u * H * 32u * 2u, kHidBytes = 2u * H * H * 2u;
    constexpr uint32_t kChunkOut = (2u * 16u * 16u) >> 4, kLoOut = (16u * H * 2u) >> 4;
    const uint32_t a = tg + col_ah<H>();
    if (type < NH) {                                 // hidden layer (type >= 1 here)
        const uint64_t b = b_first + (uint64_t)((kFirstBytes >> 4) + (uint32_t)(type - 1) * (kHidBytes >> 4));
        if (PART == 1) {
            mma_ts<0>(tg + kColDu, tg + kColAu, b, idt);
            mma_ts<0>(tg + kColDz, a, b, idz);
            mma_ts<1>(tg + kColDz, a, b + kLoHid, idz);
        } else {
            mma_ts<1>(tg + kColDu, tg + kColAu + 8, b + kChunk, idt);
            mma_ts<1>(tg + kColDz, a + 8, b + kChunk, idz);
            mma_ts<1>(tg + kColDz, a + 8, b + kLoHid + kChunk, idz);
            mma_ts<0>(tg + kColDv, tg + kColAv, b, idt);
            mma_ts<1>(tg + kColDv, tg + kColAv + 8, b + kChunk, idt);
        }
    } else {                                         // output layer, N = 16, fp32 accumulators
        const uint64_t b = b_out;
        if (PART == 1) {
            mma_ts<0>(tg + kColDu, tg + kColAu, b, ido);
            mma_ts<0>(tg + kColDz, a, b, ido);
            mma_ts<1>(tg + kColDz, a, b + kLoOut, ido);
        } else {
            mma_ts<1>(tg + kColDu, tg + kColAu + 8, b + kChunkOut, ido);
            mma_ts<1>(tg + kColDz, a + 8, b + kChunkOut, ido);
            mma_ts<1>(tg + kColDz, a + 8, b + kLoOut + kChunkOut, ido);
            mma_ts<0>(tg + kColDv, tg + kColAv, b, ido);
            mma_ts<1>(tg + kColDv, tg + kColAv + 8, b + kChunkOut, ido);
        }
    }
    if (PART == 2) tc_commit(bar);
}
// named barriers of the issuer-warp builds: 1 + g = "first half stored" (mid), 5 + g = "pass stored" (end); 160 = the group's
// 128 worker threads (bar.arrive) + its issuer warp (bar.sync)
__device__ __forceinline__ void issuer_arrive(int id) { asm volatile("bar.arrive %0, 160;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void issuer_sync(int id) { asm volatile("bar.sync %0, 160;" ::"r"(id) : "memory"); }

template <bool TANGENTS, int H>
__device__ __forceinline__ void issue_round(int type, int NH, uint32_t tg, uint64_t b_first, uint64_t b_out, uint64_t av_desc,
                                            uint32_t bar) {
    constexpr uint32_t idz = make_idesc(H, true), idt = make_idesc(H, false), ido = make_idesc(16, true);
    // descriptor address units are 16 bytes.  One K chunk (16 halves = two core-matrix columns) of an [N x K] image
    // is 2 * N * 16 bytes; the LO image follows the HI image.
    constexpr uint32_t kChunk = (2u * H * 16u) >> 4;                // K-chunk stride of an N = H operand
    constexpr uint32_t kLoFirst = (H * 32u * 2u) >> 4;              // first layer: [H x 32] halves
    constexpr uint32_t kLoHid = (H * H * 2u) >> 4;                  // hidden layer: [H x H] halves
    constexpr uint32_t kFirstBytes = 2u * H * 32u * 2u, kHidBytes = 2u * H * H * 2u;   // hi + lo
    constexpr uint32_t kChunkOut = (2u * 16u * 16u) >> 4, kLoOut = (16u * H * 2u) >> 4;
    static_assert(!TANGENTS || H == 32, "tangent rounds exist for the 32-wide sampler nets only");
    if (type < NH) {
        // layer images follow each other: bump the descriptor's 16-byte-granular address field (no carry: smem < 256 KB)
        const uint64_t b = (type == 0) ? b_first
                                       : b_first + (uint64_t)((kFirstBytes >> 4) + (uint32_t)(type - 1) * (kHidBytes >> 4));
        const uint32_t a = tg + col_ah<H>();
        const uint32_t lo = (type == 0) ? kLoFirst : kLoHid;
        if (TANGENTS && !kSmemAv) {
            mma_ts<0>(tg + kColDu, tg + kColAu, b, idt);
            if (type != 0) mma_ts<1>(tg + kColDu, tg + kColAu + 8, b + kChunk, idt);
        }
        if (H == 32) {
            mma_ts<0>(tg + kColDz, a, b, idz);
            mma_ts<1>(tg + kColDz, a + 8, b + kChunk, idz);
            mma_ts<1>(tg + kColDz, a, b + lo, idz);
            mma_ts<1>(tg + kColDz, a + 8, b + lo + kChunk, idz);
        } else {
            mma_ts<0>(tg + kColDz, a, b, idz);
            mma_ts<1>(tg + kColDz, a + 8, b + kChunk, idz);
            if (type != 0) {
#pragma unroll
                for (int c = 2; c < H / 16; ++c) mma_ts<1>(tg + kColDz, a + 8 * c, b + c * kChunk, idz);
            }
            mma_ts<1>(tg + kColDz, a, b + lo, idz);
            mma_ts<1>(tg + kColDz, a + 8, b + lo + kChunk, idz);
            if (type != 0) {
#pragma unroll
                for (int c = 2; c < H / 16; ++c) mma_ts<1>(tg + kColDz, a + 8 * c, b + lo + c * kChunk, idz);
            }
        }
        if (TANGENTS && kSmemAv) {                  // z has consumed A_h: D_u may now overwrite it
            mma_ts<0>(tg + kColDu, tg + kColAu, b, idt);
            if (type != 0) mma_ts<1>(tg + kColDu, tg + kColAu + 8, b + kChunk, idt);
            mma_ss<0>(tg + kColDv, av_desc, b, idt);
            if (type != 0) mma_ss<1>(tg + kColDv, av_desc + ((2u * 2048u) >> 4), b + kChunk, idt);
        } else if (TANGENTS) {
            mma_ts<0>(tg + kColDv, tg + kColAv, b, idt);
            if (type != 0) mma_ts<1>(tg + kColDv, tg + kColAv + 8, b + kChunk, idt);
        }
    } else {                                        // output layer, N = 16
        const uint64_t b = b_out;
        const uint32_t ah = tg + col_ah<H>();
        if (TANGENTS && !kSmemAv) {
            mma_ts<0>(tg + kColDu, tg + kColAu, b, ido);
            mma_ts<1>(tg + kColDu, tg + kColAu + 8, b + kChunkOut, ido);
        }
        mma_ts<0>(tg + kColDz, ah, b, ido);
#pragma unroll
        for (int c = 1; c < H / 16; ++c) mma_ts<1>(tg + kColDz, ah + 8 * c, b + c * kChunkOut, ido);
#pragma unroll
        for (int c = 0; c < H / 16; ++c) mma_ts<1>(tg + kColDz, ah + 8 * c, b + kLoOut + c * kChunkOut, ido);
        if (TANGENTS && kSmemAv) {
            mma_ts<0>(tg + kColDu, tg + kColAu, b, ido);
            mma_ts<1>(tg + kColDu, tg + kColAu + 8, b + kChunkOut, ido);
            mma_ss<0>(tg + kColDv, av_desc, b, ido);
            mma_ss<1>(tg + kColDv, av_desc + ((2u * 2048u) >> 4), b + kChunkOut, ido);
        } else if (TANGENTS) {
            mma_ts<0>(tg + kColDv, tg + kColAv, b, ido);
            mma_ts<1>(tg + kColDv, tg + kColAv + 8, b + kChunkOut, ido);
        }
    }
    tc_commit(bar);
}

// Hand the freshly written A operands to the tensor core: all 128 threads' TMEM stores must have landed
// before one elected lane issues the round.
template <bool TANGENTS, int H>
__device__ __forceinline__ void publish_and_issue(int g, int q, int type, int NH, uint32_t tg_mma, uint64_t b_hid0,
                                                  uint64_t b_out, uint64_t av_desc, uint32_t bar, int trace_r = 0) {
    if (kSmemAv && TANGENTS) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared operands -> tensor core
    tc_wait_st();
    tc_fence_before();
    TC_TRACE(4);
    if (kIssuer && TANGENTS && H == 32) { issuer_arrive(5 + g); return; }      // the group's issuer warp takes it from here
    group_sync(g);
    TC_TRACE(5);
    if (q == 0) {
        if (elect_one()) { tc_fence_after(); issue_round<TANGENTS, H>(type, NH, tg_mma, b_hid0, b_out, av_desc, bar); }
        TC_TRACE(6);
    }
}

// First-layer operand of one Euler step for this thread's row: A1 (over A_h) = state hi/lo parts in columns 0..3, PE5(wi)
// from the tile's shared-memory record in 4..14, zero in 15; tangent seeds d(input row)/d(state) -> A_u, A_v (K chunk 0).
template <int DOMAIN, bool TANGENTS, int H>
__device__ __forceinline__ void build_first_operand(const float (*f)[kTile], int row, float x0, float x1, float alpha,
                                                    uint32_t tg, uint32_t av_row) {
    uint32_t a1[16];
#pragma unroll
    for (int j = 0; j < 11; ++j) a1[4 + j] = __float_as_uint(f[kFPe + j][row]);
    a1[15] = 0u;
    float s0, s1, s2v, s3;
    if (DOMAIN == kDisk) { s0 = x0; s1 = x1; s2v = alpha; s3 = 0.0f; }
    else { __sincosf(x1, &s1, &s2v); s0 = x0; s3 = alpha; }
    const uint32_t c01 = pack_h2(s0, s1), c23 = pack_h2(s2v, s3);
    const uint32_t l01 = pack_h2(s0 - h2_lo(c01), s1 - h2_hi(c01));
    const uint32_t l23 = pack_h2(s2v - h2_lo(c23), s3 - h2_hi(c23));
    uint32_t eu[8] = {0x00003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};                    // d/dx0 (d/dtheta): k = 0
    uint32_t ev[8] = {0x3C000000u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};                    // disk d/dx1: k = 1
    if (DOMAIN == kDisk) {
        // k: x0_hi x1_hi | a_hi a_lo | x0_lo x1_lo | 0 0
        a1[0] = c01; a1[1] = (c23 & 0xffffu) | (l23 << 16); a1[2] = l01; a1[3] = 0u;
    } else {
        // k: th_hi sin_hi | cos_hi a_hi | th_lo sin_lo | cos_lo a_lo
        a1[0] = c01; a1[1] = c23; a1[2] = l01; a1[3] = l23;
        // d/dphi: d(sin) = cos on k = 1, d(cos) = -sin on k = 2 (hi parts), lo parts on k = 5, 6 (the
        // weight image repeats W1[:,1], W1[:,2] there)
        ev[0] = c23 << 16; ev[1] = (c01 >> 16) ^ 0x8000u; ev[2] = l23 << 16; ev[3] = (l01 >> 16) ^ 0x8000u;
    }
    tmem_st8(tg + col_ah<H>(), a1);
    tmem_st8(tg + col_ah<H>() + 8, a1 + 8);
    if (TANGENTS) {
        tmem_st8(tg + kColAu, eu);
        if (kSmemAv) smem_st_a16(av_row, 0, ev); else tmem_st8(tg + kColAv, ev);
    }
}

// One activation pass over this thread's row of a tile whose round has completed: tcgen05.ld z (fp32), du, dv (packed
// fp16) -> h = silu(z), u = s2 du, v = s2 dv -> fp16 operand words of the next round (tcgen05.st).  The second half of the
// accumulators streams in under the first half's math.
template <bool TANGENTS, int ACT, int H>
__device__ __forceinline__ void activation_pass(uint32_t tg, uint32_t av_row, int trace_r = 0, int g = 0) {
    if (H != 32) {
        // wide forward-only round: H / 16 chunks of 16 neurons, the next chunk's load in flight
        float zc[2][16];
        uint32_t ph[8];
        tmem_ld16(tg + kColDz, zc[0]);
#pragma unroll
        for (int c = 0; c < H / 16; ++c) {
            tc_wait_ld();
            if (c + 1 < H / 16) tmem_ld16(tg + kColDz + 16 * (c + 1), zc[(c + 1) & 1]);
            activate16<false, ACT>(zc[c & 1], nullptr, nullptr, ph, nullptr, nullptr);
            tmem_st8(tg + col_ah<H>() + 8 * c, ph);
        }
        return;
    }
    if (BSDFDIFF_TC_PIPE4 && TANGENTS && ACT == 1 && !kSmemAv) {
        float z[4][8], t[2][8];
        uint32_t du[4][4], dv[4][4], ph[4], pu[4], pv[4];
        auto ld = [&](int c) {
            tmem_ld8(tg + kColDz + 8 * c, z[c]);
            tmem_ld4_pack16(tg + kColDu + 8 * c, du[c]);
            tmem_ld4_pack16(tg + kColDv + 8 * c, dv[c]);
        };
        auto st = [&](int c) {
            tmem_st4(tg + col_ah<H>() + 4 * c, ph[0], ph[1], ph[2], ph[3]);
            tmem_st4(tg + kColAu + 4 * c, pu[0], pu[1], pu[2], pu[3]);
            tmem_st4(tg + kColAv + 4 * c, pv[0], pv[1], pv[2], pv[3]);
        };
        // A_u (columns kColAu + 0..15) overlaps D_v's FIRST 16 columns only = chunks 0 and 1, which are in registers after the
        // first wait::ld: every store below is safe
        ld(0); ld(1);
        tc_wait_ld();
        ld(2);
        tanh8(z[0], t[0]);
        tanh8(z[1], t[1]); tail8(z[0], t[0], du[0], dv[0], ph, pu, pv); st(0);
        tc_wait_ld();
        ld(3);
        tanh8(z[2], t[0]); tail8(z[1], t[1], du[1], dv[1], ph, pu, pv); st(1);
        tc_wait_ld();
        tanh8(z[3], t[1]); tail8(z[2], t[0], du[2], dv[2], ph, pu, pv); st(2);
        tail8(z[3], t[1], du[3], dv[3], ph, pu, pv); st(3);
        return;
    }
    float za[16], zb[16];
    uint32_t ua[TANGENTS ? 8 : 1], va[TANGENTS ? 8 : 1], ub[TANGENTS ? 8 : 1], vb[TANGENTS ? 8 : 1];
    tmem_ld16(tg + kColDz, za);
    if (TANGENTS) { tmem_ld8_pack16(tg + kColDu, ua); tmem_ld8_pack16(tg + kColDv, va); }
    tc_wait_ld();
    TC_TRACE(1);
    tmem_ld16(tg + kColDz + 16, zb);           // second half streams in under the first half's math
    if (TANGENTS) { tmem_ld8_pack16(tg + kColDu + 16, ub); tmem_ld8_pack16(tg + kColDv + 16, vb); }
    uint32_t ph[8], pu[8], pv[8];
    activate16<TANGENTS, ACT>(za, ua, va, ph, pu, pv);
    tmem_st8(tg + col_ah<H>(), ph);
    if (TANGENTS) {
        tmem_st8(tg + kColAu, pu);
        if (kSmemAv) smem_st_a16(av_row, 0, pv); else tmem_st8(tg + kColAv, pv);
    }
    tc_wait_ld();
    TC_TRACE(2);
    if (BSDFDIFF_TC_ISSUER == 1 && TANGENTS) {     // first half stored, every accumulator of the round read: chunk 0 may go
        tc_wait_st();
        tc_fence_before();
        issuer_arrive(1 + g);
    }
    activate16<TANGENTS, ACT>(zb, ub, vb, ph, pu, pv);
    tmem_st8(tg + col_ah<H>() + 8, ph);
    if (TANGENTS) {
        tmem_st8(tg + kColAu + 8, pu);
        if (kSmemAv) smem_st_a16(av_row, 2, pv); else tmem_st8(tg + kColAv + 8, pv);
    }
    TC_TRACE(3);
}

// Duo workers: hand a pipe's freshly written operands to the tensor core without making anybody wait.
//   mode 2: every warp bumps the pipe's shared-memory arrival counter (acq_rel) after its stores have landed; the warp
//           that observes the fourth arrival of the round issues the MMAs.  Arrivals of round r+1 can only start after
//           round r's MMAs have completed (every warp waits on d_ready first), so the counter never mixes rounds.
//   mode 1: a rotating warp (round number mod 4) blocks in bar.sync, the other three bar.arrive and go on.
//   mode 0: all four warps bar.sync; the group's first warp issues.
template <bool TANGENTS, int H>
__device__ __forceinline__ void publish_and_issue_duo(TcSmem& S, int pipe, int q, int lane, uint32_t& rounds, int type,
                                                      int NH, uint32_t tg_mma, uint64_t b_hid0, uint64_t b_out,
                                                      uint32_t bar) {
    tc_wait_st();
    tc_fence_before();
    bool mine;
    if (BSDFDIFF_TC_ISSUE == 2) {
        uint32_t old = 0;
        __syncwarp();
        if (lane == 0)
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&S.arrive[pipe])) : "memory");
        old = __shfl_sync(0xffffffffu, old, 0);
        mine = (old & 3u) == 3u;
    } else if (BSDFDIFF_TC_ISSUE == 1) {
        mine = (int)(rounds & 3u) == q;
        if (mine) asm volatile("bar.sync %0, %1;" ::"r"(pipe + 1), "r"(128) : "memory");
        else asm volatile("bar.arrive %0, %1;" ::"r"(pipe + 1), "r"(128) : "memory");
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(pipe + 1), "r"(128) : "memory");
        mine = (q == 0);
    }
    ++rounds;
    if (mine) {
        if (elect_one()) { tc_fence_after(); issue_round<TANGENTS, H>(type, NH, tg_mma, b_hid0, b_out, 0ull, bar); }
    }
}

// Per-thread state of one of the two tiles a duo worker thread owns.
struct Pipe {
    float x0, x1, R, p0, theta_o, x1_start;
    CondTrack cond;
    long long k;            // index of the tile in this CTA's tile list (k % kGroups == pipe)
    int sl;                 // ring slot of tile k
    uint32_t use;           // ring lap of tile k
    uint32_t pd;            // phase parity of the pipe's d_ready barrier
    uint32_t rounds;        // rounds published so far (issue mode 1)
    int t;                  // Euler step
    int phase;              // -1: take the next tile; 0..NH-1: activation pass of that round; NH: output pass
    bool alive;
};

// Raw per-query inputs of one tile row, loaded one tile ahead by the producer so the global-load latency is
// never on anybody's critical path.
struct RawIn {
    float wa, wb, wc;        // wi: (w0, w1, -) raw epilogue, or the 3 local-frame components
    float oa, ob, oc;        // wo (pdf mode)
    float r0, r1;            // replayed base sample (or, with r2, the renderer's uniforms)
    float r2;
};
template <int MODE>
__device__ __forceinline__ void raw_load(const FlowParams& P, long long i, RawIn& r) {
    const long long qi = (P.wi_repeat > 1) ? i / P.wi_repeat : i;
    if (P.epilogue == kEpiRaw) {
        const float2 w = reinterpret_cast<const float2*>(P.wi)[qi];
        r.wa = w.x; r.wb = w.y; r.wc = 1.0f;
    } else {
        const float* w = P.wi + qi * P.wi_l.rs;
        r.wa = w[0]; r.wb = w[P.wi_l.c1]; r.wc = w[P.wi_l.c2];
    }
    if (MODE == kModePdf) {
        if (P.epilogue == kEpiRaw) {
            const float2 w = reinterpret_cast<const float2*>(P.wo)[i];
            r.oa = w.x; r.ob = w.y; r.oc = 1.0f;
        } else {
            const float* w = P.wo + i * P.wo_l.rs;
            r.oa = w[0]; r.ob = w[P.wo_l.c1]; r.oc = w[P.wo_l.c2];
        }
    } else if (P.x0) {
        const float2 t = reinterpret_cast<const float2*>(P.x0)[i];
        r.r0 = t.x; r.r1 = t.y;
    } else if (P.u_noise) {
        r.r0 = P.u_noise[3 * i]; r.r1 = P.u_noise[3 * i + 1]; r.r2 = P.u_noise[3 * i + 2];
    }
}

// ------------------------------------------------------------------------------------------------
// The kernel.  Warp roles (one persistent CTA per SM, kTcThreads = 640 threads = 20 warps in the shipped build):
//   warps 0..15   four worker groups of 4 warps (kGroups = 4); thread <-> TMEM lane <-> query row of the group's current
//                 128-query tile, group g owns TMEM columns [128 g, 128 g + 128).  They run nothing but the T-step flow:
//                 activation math between tensor-core round trips; warp 4 g issues group g's MMAs.
//   warps 16..19  the producer group; thread <-> query row.  It runs up to two tiles ahead of the workers and does all
//                 per-query set-up that needs no tensor core: wi (and wo / replayed noise / renderer uniforms) loads,
//                 PE5(wi), the base net, the Philox base sample and its density.  Records go through a 6-slot shared-memory
//                 ring (full/empty mbarriers).  This work is latency-bound (global loads, 10 serial Philox rounds, serial
//                 FMA chains); on its own warps it fills the issue slots the workers leave idle while they wait for MMAs
//                 instead of adding ~1/3 to every tile's critical path (profiles/r1d vs r1h).
//   MULTI         one wavefront, several materials (multi.cu): the tile list is the plan's virtual-tile table, every group
//                 keeps its own weight image in shared memory and re-stages it (one cp.async.bulk) when its next tile's
//                 material differs; the producer re-stages the base net and passes each row's wavefront index through
//                 the ring.
// (tuning builds: BSDFDIFF_TC_GROUPS = 1..5, BSDFDIFF_TC_NOALIAS = the 144-column map with three groups, BSDFDIFF_TC_DUO =
//  two groups x two tiles per thread)
// ------------------------------------------------------------------------------------------------
template <int DOMAIN, int MODE, int ACT, int H, bool MULTI>
__global__ void __launch_bounds__(kTcThreads, 1) flow_tc_kernel(const FlowParams P) {
    constexpr bool TANGENTS = (MODE != kModeForward);
    static_assert(!MULTI || (!kDuo && !kSmemAv), "multi-material launches use the one-tile-per-thread worker structure");
    static_assert(H == 32 || (H == 64 && !TANGENTS), "64-wide nets: forward-only (reflow teacher) rounds");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
    // warp index through a shuffle so the compiler can prove it warp-uniform: TMEM addresses and MMA
    // descriptors then live in uniform registers (no per-MMA R2UR/ELECT waterfall)
    const int warp = (int)__reduce_max_sync(0xffffffffu, (unsigned)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
    const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(MULTI ? P.flows[0] : P.flow);   // multi: all blobs share one shape
    const int NH = P.n_hidden;
    const uint32_t img_bytes = (hdr->f16_bytes + 127u) & ~127u;                  // multi: one weight image per group

    // ---- one-time setup ---------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int g = 0; g < kGroups; ++g) { mbar_init(smem_u32(&S.d_ready[g]), 1); S.arrive[g] = 0u; }
        for (int sl = 0; sl < kSlots; ++sl) {
            mbar_init(smem_u32(&S.full[sl]), kProducerWarps);      // one arrival per producer warp
            mbar_init(smem_u32(&S.empty[sl]), 4);                  // one arrival per warp of the consuming group
        }
        for (int g = 0; g < kGroups; ++g) mbar_init(smem_u32(&S.w_bar[g]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!MULTI) {
            const uint32_t f16_bytes = hdr->f16_bytes;
            mbar_expect_tx(smem_u32(&S.w_bar[0]), f16_bytes);
            tma_bulk_g2s(smem_u32(S.w16), P.flow + hdr->off_f16, f16_bytes, smem_u32(&S.w_bar[0]));
        }
    }
    if (!MULTI && P.base) {
        for (int i = threadIdx.x; i < kPE3 * 16; i += kTcThreads) {          // w1t[k][j] = W1[j][k]
            const int k = i >> 4, j = i & 15;
            S.bw1t[i] = P.base[j * kPE3 + k];
        }
        if (threadIdx.x < 16) S.bb1[threadIdx.x] = P.base[224 + threadIdx.x];
        if (threadIdx.x < 64) {                                              // wot[j][c] = Wo[c][j]
            const int j = threadIdx.x >> 2, c = threadIdx.x & 3;
            S.bwot[threadIdx.x] = P.base[240 + c * 16 + j];
        }
        if (threadIdx.x < 4) S.bbo[threadIdx.x] = P.base[304 + threadIdx.x];
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&S.tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, S.tmem_base);

    const long long n_tiles = MULTI ? (long long)*P.n_tiles_dev : (P.n + kTile - 1) / kTile;
    // tile list of this CTA: blockIdx.x, +gridDim.x, ...; entry k goes to group k % kGroups through slot k % kSlots
    const long long my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (kIssuer && warp >= kWGroups * 4 + kProducerWarps) {
        // =========================== MMA issuer of group g (tuning build) ================================
        if (TANGENTS && H == 32) {
            const int g = warp - (kWGroups * 4 + kProducerWarps);
            const uint32_t tg_mma = tmem_base + g * kColsPerGroup;
            const uint32_t bar_d = smem_u32(&S.d_ready[g]);
            const uint32_t w_base = smem_u32(S.w16) + (MULTI ? (uint32_t)g * img_bytes : 0u);
            const uint64_t b_hid0 = make_b_desc(w_base, H * 16, 128);
            uint64_t b_out = make_b_desc(w_base + 2u * H * 32u * 2u + (uint32_t)(NH - 1) * (2u * H * H * 2u), 256, 128);
            asm volatile("" : "+l"(b_out));
            int mat_staged = -1;
            uint32_t wpar = 0;
            if (!MULTI) mbar_wait(smem_u32(&S.w_bar[0]), 0);
#pragma unroll 1
            for (long long k = g; k < my_tiles; k += kGroups) {
#pragma unroll 1
                for (int t = 0; t < P.T; ++t) {
                    issuer_sync(5 + g);                       // the step's first-layer operand is stored
                    if (MULTI && t == 0) {
                        // (every MMA of the group's previous tile has completed: its workers waited for the last round before
                        // they built this operand)
                        const int mat = P.tiles[blockIdx.x + k * gridDim.x].x;
                        if (mat != mat_staged) {
                            if (elect_one()) {
                                const unsigned char* blob = P.flows[mat];
                                const PackedHeader* h2 = reinterpret_cast<const PackedHeader*>(blob);
                                mbar_expect_tx(smem_u32(&S.w_bar[g]), h2->f16_bytes);
                                tma_bulk_g2s(w_base, blob + h2->off_f16, h2->f16_bytes, smem_u32(&S.w_bar[g]));
                            }
                            __syncwarp();
                            mbar_wait(smem_u32(&S.w_bar[g]), wpar); wpar ^= 1u;
                            mat_staged = mat;
                        }
                    }
                    if (elect_one()) { tc_fence_after(); issue_round<TANGENTS, H>(0, NH, tg_mma, b_hid0, b_out, 0ull, bar_d); }
                    __syncwarp();
#pragma unroll 1
                    for (int l = 0; l < NH; ++l) {
                        if (BSDFDIFF_TC_ISSUER == 1) {            // split issue: chunk 0 of u and z under the second half's math
                            issuer_sync(1 + g);
                            if (elect_one()) { tc_fence_after(); issue_round_split<1>(l + 1, NH, tg_mma, b_hid0, b_out, bar_d); }
                            __syncwarp();
                            issuer_sync(5 + g);
                            if (elect_one()) { tc_fence_after(); issue_round_split<2>(l + 1, NH, tg_mma, b_hid0, b_out, bar_d); }
                            __syncwarp();
                        } else {                                  // == 2: dedicated issuer, whole round at the end of the pass
                            issuer_sync(5 + g);
                            if (elect_one()) { tc_fence_after(); issue_round<TANGENTS, H>(l + 1, NH, tg_mma, b_hid0, b_out, 0ull, bar_d); }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    } else if (warp >= kWGroups * 4) {
        // =========================== producer ==========================================================
        const int row = (warp - kWGroups * 4) * 32 + lane;
        RawIn cur, nxt;
        cur.wa = cur.wb = cur.wc = cur.oa = cur.ob = cur.oc = cur.r0 = cur.r1 = cur.r2 = 0.0f;
        nxt = cur;
        int mat_n = 0;                                      // material of the tile index_of looked at last
        auto index_of = [&](long long k, bool& valid) {
            if (MULTI) {
                const int4 ti = P.tiles[blockIdx.x + k * gridDim.x];
                mat_n = ti.x;
                valid = row < ti.z;
                return (long long)P.perm[ti.y + (valid ? row : 0)];   // padding rows recompute the tile's first row, never store
            }
            const long long i_raw = (blockIdx.x + k * gridDim.x) * kTile + row;
            valid = i_raw < P.n;
            return valid ? i_raw : (P.n - 1);               // tail rows recompute the last query, never store
        };
        bool valid = false, valid_n = false;
        long long i = 0, i_n = 0;
        int mat = -1, mat_staged = -1;
        int sl = 0;                                         // ring position of tile k: k % kSlots, lap k / kSlots
        uint32_t use = 0;
        if (my_tiles > 0) { i = index_of(0, valid); mat = mat_n; raw_load<MODE>(P, i, cur); }
#pragma unroll 1
        for (long long k = 0; k < my_tiles; ++k) {
            const int mat_cur = mat;
            if (k + 1 < my_tiles) { i_n = index_of(k + 1, valid_n); mat = mat_n; raw_load<MODE>(P, i_n, nxt); }
            if (MULTI && mat_cur != mat_staged) {
                // this tile's material differs from the staged base net: the 128 producer threads re-stage it
                // (only they read it; named barrier 15 keeps the four producer warps in step)
                asm volatile("bar.sync 15, 128;" ::: "memory");
                const float* bsrc = P.bases[mat_cur];
                const int t = row;
                for (int q2 = t; q2 < kPE3 * 16; q2 += 128) S.bw1t[q2] = bsrc[(q2 & 15) * kPE3 + (q2 >> 4)];
                if (t < 16) S.bb1[t] = bsrc[224 + t];
                if (t < 64) S.bwot[t] = bsrc[240 + (t & 3) * 16 + (t >> 2)];
                if (t < 4) S.bbo[t] = bsrc[304 + t];
                asm volatile("bar.sync 15, 128;" ::: "memory");
                mat_staged = mat_cur;
            }
            // conditioning in domain coordinates (brdf_measured_disk.py:66-67, brdf_measured_spherical.py:76-77)
            float w0, w1, wiz = cur.wc;
            if (P.epilogue == kEpiRaw || P.epilogue == kEpiDisk) { w0 = cur.wa; w1 = cur.wb; }
            else cart_to_spher(cur.wa, cur.wb, cur.wc, w0, w1);
            float e[kPE5];
            pe5_fast(w0, w1, e);
            uint32_t c[11];
#pragma unroll
            for (int j = 0; j < 11; ++j) c[j] = pack_h2(e[2 * j], e[2 * j + 1]);
            float bp[4] = {0.f, 0.f, 0.f, 0.f};
            if (MULTI || P.base) base_eval_smem(S, e, bp);      // PE3 is a prefix of PE5
            float x0, x1, p0 = 1.0f;
            if (MODE == kModePdf) {
                if (P.epilogue == kEpiRaw || P.epilogue == kEpiDisk) { x0 = cur.oa; x1 = cur.ob; }   // brdf_measured_disk.py:118-120
                else cart_to_spher(cur.oa, cur.ob, cur.oc, x0, x1);                                  // brdf_measured_spherical.py:131-133
            } else {
                if (P.x0) { x0 = cur.r0; x1 = cur.r1; }
                else if (P.u_noise) { const float un[3] = {cur.r0, cur.r1, cur.r2}; base_draw_fast<DOMAIN>(bp, 0ull, 0ull, 0, x0, x1, un); }
                else base_draw_fast<DOMAIN>(bp, P.seed, P.offset, P.first_index + i, x0, x1);
                if (P.out_x0 && valid) reinterpret_cast<float2*>(P.out_x0)[i] = make_float2(x0, x1);
                if (MODE == kModeSample) p0 = __expf(base_logprob_fast<DOMAIN>(bp, x0, x1));
            }

            mbar_wait<BSDFDIFF_TC_BACKOFF_NS>(smem_u32(&S.empty[sl]), (use & 1u) ^ 1u);   // the group that used this slot last has read it
            float (*f)[kTile] = S.slot[sl];
#pragma unroll
            for (int j = 0; j < 11; ++j) f[kFPe + j][row] = __uint_as_float(c[j]);
            f[kFX][row] = x0; f[kFX + 1][row] = x1;
            if (MODE == kModeSample) f[kFP0][row] = p0;
            if (MODE == kModePdf) {
#pragma unroll
                for (int j = 0; j < 4; ++j) f[kFBp + j][row] = bp[j];
                f[kFWiz][row] = wiz;
                f[kFWo][row] = cur.oa; f[kFWo + 1][row] = cur.ob; f[kFWo + 2][row] = cur.oc;
            }
            if (MULTI) f[kFIdx][row] = __int_as_float(valid ? (int)i : -1);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S.full[sl]));      // release: the stores above are visible to the waiters
            cur = nxt; i = i_n; valid = valid_n;
            if (++sl == kSlots) { sl = 0; ++use; }
        }
    } else if (kDuo) {
        // =========================== workers, two tiles per thread ====================================
        const int g = warp >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t w_base = smem_u32(S.w16);
        const uint64_t b_hid0 = make_b_desc(w_base, H * 16, 128);                // first / hidden layers: N = H rows
        uint64_t b_out = make_b_desc(w_base + 2u * H * 32u * 2u + (uint32_t)(NH - 1) * (2u * H * H * 2u), 256, 128);   // N = 16
        asm volatile("" : "+l"(b_out));
        const float inv_t = (float)(1.0 / (double)P.T);
        const float step = (MODE == kModePdf) ? -inv_t : inv_t;
        mbar_wait(smem_u32(&S.w_bar[0]), 0);                                    // weights have landed in smem (any warp may issue)

        // one turn of tile slot SLOT (0 | 1) of this thread: pipe = g + 2 SLOT (both groups get a tile before either gets two)
        auto turn = [&](auto slot_tag, Pipe& p) {
            constexpr int SLOT = decltype(slot_tag)::value;
            const int pipe = g + kWGroups * SLOT;
            const uint32_t tg_mma = tmem_base + pipe * kColsPerGroup;            // lane field 0: MMA operand addresses
            const uint32_t tg = tg_mma + ((uint32_t)(q * 32) << 16);             // this warp's 32-lane window
            const uint32_t bar_d = smem_u32(&S.d_ready[pipe]);
            bool build = false;
            if (p.phase >= 0) {
                mbar_wait(bar_d, p.pd); p.pd ^= 1u;
                tc_fence_after();
                if (p.phase < NH) {
                    // ---- activation pass of round `phase`; the next round goes to the tensor core ----
                    activation_pass<TANGENTS, ACT, H>(tg, 0u);
                    publish_and_issue_duo<TANGENTS, H>(S, pipe, q, lane, p.rounds, p.phase + 1, NH, tg_mma, b_hid0, b_out, bar_d);
                    ++p.phase;
                    return;
                }
                // ---- output pass: d, dd/dx0, dd/dx1 -> determinant, Euler update ----
                float d0, d1, du0 = 0.f, du1 = 0.f, dv0 = 0.f, dv1 = 0.f;
                tmem_ld2(tg + kColDz, d0, d1);
                if (TANGENTS) { tmem_ld2(tg + kColDu, du0, du1); tmem_ld2(tg + kColDv, dv0, dv1); }
                tc_wait_ld();
                if (TANGENTS) {
                    const float j00 = __fmaf_rn(step, du0, 1.0f), j01 = __fmul_rn(step, dv0);
                    const float j10 = __fmul_rn(step, du1), j11 = __fmaf_rn(step, dv1, 1.0f);
                    const float det = __fmaf_rn(j00, j11, -__fmul_rn(j01, j10));
                    p.R = (MODE == kModePdf) ? __fmul_rn(p.R, det) : __fdividef(p.R, det);
                    p.cond.step(j00, j01, j10, j11, det);
                }
                p.x0 = fmaf(step, d0, p.x0);
                p.x1 = fmaf(step, d1, p.x1);
                if (++p.t < P.T) {
                    build = true;
                } else {
                    // ---- tile done: epilogue, hand the ring slot back, move to this pipe's next tile ----
                    const long long i_raw = (blockIdx.x + p.k * gridDim.x) * kTile + row;
                    const bool valid = i_raw < P.n;
                    const long long i = valid ? i_raw : (P.n - 1);
                    const float (*f)[kTile] = S.slot[p.sl];
                    if (MODE == kModeSample) {
                        if (valid) store_sample<true>(P, i, p.x0, p.x1, p.p0 * p.R);
                        if (P.fix_thr > 0.0f) flag_for_fixup(P, valid && p.cond.weight() < P.fix_thr, i, -1, p.theta_o, p.x1_start);
                    } else if (MODE == kModePdf) {
                        float bp[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) bp[j] = f[kFBp + j][row];
                        const float wiz = f[kFWiz][row], wox = f[kFWo][row], woy = f[kFWo + 1][row], woz = f[kFWo + 2][row];
                        if (valid)
                            store_pdf<true>(P, i, __expf(base_logprob_fast<DOMAIN>(bp, p.x0, p.x1)) * p.R, wiz, wox, woy, woz,
                                            p.theta_o);
                        if (P.fix_thr > 0.0f) {
                            const float kappa = (DOMAIN == kDisk) ? 0.0f : softplus_fast(bp[3]) + 1e-3f;
                            const float gn = base_grad_norm(DOMAIN, bp, kappa, p.x0, p.x1);
                            flag_for_fixup(P, valid && p.cond.weight() * fminf(1.0f, __fdividef(25.0f, gn)) < P.fix_thr, i);
                        }
                    } else {
                        if (valid) reinterpret_cast<float2*>(P.out_dir)[i] = make_float2(p.x0, p.x1);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.empty[p.sl]));        // every read of the record is done
                    p.k += kGroups;
                    p.sl += kGroups;
                    if (p.sl >= kSlots) { p.sl -= kSlots; ++p.use; }
                    p.phase = -1;
                    p.alive = p.k < my_tiles;
                    if (!p.alive) return;
                }
            }
            if (p.phase < 0) {
                // ---- take this row's record of the pipe's next tile from the producer ----
                mbar_wait(smem_u32(&S.full[p.sl]), p.use & 1u);
                const float (*f)[kTile] = S.slot[p.sl];
                p.x0 = f[kFX][row]; p.x1 = f[kFX + 1][row];
                p.p0 = (MODE == kModeSample) ? f[kFP0][row] : 1.0f;
                p.theta_o = p.x0; p.x1_start = p.x1;
                p.R = 1.0f;
                p.cond.reset();
                p.t = 0;
                build = true;
            }
            if (build) {
                // ---- first-layer operand of step t, round 0 to the tensor core ----
                const float tf = (float)p.t / (float)P.T;
                const float alpha = (MODE == kModePdf) ? 1.0f - tf : tf;
                build_first_operand<DOMAIN, TANGENTS, H>(S.slot[p.sl], row, p.x0, p.x1, alpha, tg, 0u);
                publish_and_issue_duo<TANGENTS, H>(S, pipe, q, lane, p.rounds, 0, NH, tg_mma, b_hid0, b_out, bar_d);
                p.phase = 0;
            }
        };

        Pipe pa, pb;
        pa.k = g;            pa.sl = g;            pa.alive = pa.k < my_tiles;
        pb.k = g + kWGroups; pb.sl = g + kWGroups; pb.alive = pb.k < my_tiles;
        pa.use = pb.use = 0u; pa.pd = pb.pd = 0u; pa.rounds = pb.rounds = 0u; pa.phase = pb.phase = -1; pa.t = pb.t = 0;
        pa.x0 = pa.x1 = pb.x0 = pb.x1 = 0.0f; pa.R = pb.R = pa.p0 = pb.p0 = 1.0f; pa.theta_o = pb.theta_o = 0.0f; pa.x1_start = pb.x1_start = 0.0f;
        pa.cond.reset(); pb.cond.reset();
        int skew = BSDFDIFF_TC_SKEW;
#pragma unroll 1
        while (pa.alive || pb.alive) {
            if (pa.alive) turn(std::integral_constant<int, 0>{}, pa);
            if (pb.alive) {
                if (skew > 0 && pa.alive) --skew;
                else turn(std::integral_constant<int, 1>{}, pb);
            }
        }
    } else {
        // =========================== workers, one tile per thread (round-1 structure; A/B and NOALIAS builds) ==========
        const int g = warp >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t tg_mma = tmem_base + g * kColsPerGroup;                   // lane field 0: MMA operand addresses
        const uint32_t tg = tg_mma + ((uint32_t)(q * 32) << 16);                 // this warp's 32-lane window
        const uint32_t bar_d = smem_u32(&S.d_ready[g]);
        const uint32_t w_base = smem_u32(S.w16) + (MULTI ? (uint32_t)g * img_bytes : 0u);   // multi: this group's own weight image
        const uint64_t b_hid0 = make_b_desc(w_base, H * 16, 128);                // first / hidden layers: N = H rows
        uint64_t b_out = make_b_desc(w_base + 2u * H * 32u * 2u + (uint32_t)(NH - 1) * (2u * H * H * 2u), 256, 128);   // N = 16
        uint64_t av_desc = kSmemAv ? make_b_desc(smem_u32(S.av[kSmemAv ? g : 0]), 2048, 128) : 0ull;   // A image: LBO 2048, SBO 128
        const uint32_t av_row = smem_u32(S.av[kSmemAv ? g : 0]) + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
        asm volatile("" : "+l"(b_out));                                         // keep it in registers (no per-round rebuild)
        uint32_t pd = 0;
        int sl = g;                                         // ring position / lap of this group's current tile
        uint32_t use = 0;
        const float inv_t = (float)(1.0 / (double)P.T);
        const float step = (MODE == kModePdf) ? -inv_t : inv_t;
        if (!(kIssuer && TANGENTS && H == 32) && !MULTI && q == 0) mbar_wait(smem_u32(&S.w_bar[0]), 0);  // weights have landed in smem
        int mat = 0, mat_staged = -1;                                           // multi: material of this tile / of the group's image
        uint32_t wpar = 0;
        int trace_r = 0;
        (void)trace_r;
#ifdef BSDFDIFF_TC_STAGGER_NS
        // tuning build: start the groups out of phase (group g waits g * STAGGER ns before its first tile)
        if (g > 0) asm volatile("nanosleep.u32 %0;" ::"r"((unsigned)(g * BSDFDIFF_TC_STAGGER_NS)) : "memory");
#endif

#pragma unroll 1
        for (long long k = g; k < my_tiles; k += kGroups) {
            long long i_raw = (blockIdx.x + k * gridDim.x) * kTile + row;
            bool valid = i_raw < P.n;
            if (MULTI) {
                // the tile's material: its weight image replaces the group's previous one.  Every MMA that read the old
                // image has completed (all four warps waited for the previous tile's last round), and only this warp
                // issues the group's MMAs, so it alone has to see the new image land.
                mat = P.tiles[blockIdx.x + k * gridDim.x].x;
                if (!(kIssuer && TANGENTS && H == 32) && q == 0 && mat != mat_staged) {
                    if (elect_one()) {
                        const unsigned char* blob = P.flows[mat];
                        const PackedHeader* h2 = reinterpret_cast<const PackedHeader*>(blob);
                        const uint32_t bytes = h2->f16_bytes;
                        mbar_expect_tx(smem_u32(&S.w_bar[g]), bytes);
                        tma_bulk_g2s(w_base, blob + h2->off_f16, bytes, smem_u32(&S.w_bar[g]));
                    }
                    __syncwarp();
                    mbar_wait(smem_u32(&S.w_bar[g]), wpar); wpar ^= 1u;
                }
                mat_staged = mat;
            }

            // ---- take this row's record from the producer ----
            float x0, x1, R = 1.0f, p0 = 1.0f;
            CondTrack cond;
            cond.reset();
            mbar_wait(smem_u32(&S.full[sl]), use & 1u);
            const float (*f)[kTile] = S.slot[sl];
            x0 = f[kFX][row]; x1 = f[kFX + 1][row];
            if (MODE == kModeSample) p0 = f[kFP0][row];
            if (MULTI) {                                     // wavefront row of this tile row (-1: padding row)
                const int idx = __float_as_int(f[kFIdx][row]);
                valid = idx >= 0;
                i_raw = idx;
            }
            const long long i = valid ? i_raw : (MULTI ? 0 : P.n - 1);
            const float theta_o = x0, x1_start = x1;          // start state: pdf-mode mask / the base sample a fix-up replays

#pragma unroll 1
            for (int t = 0; t < P.T; ++t) {
                const float tf = __fdividef((float)t, (float)P.T);
                const float alpha = (MODE == kModePdf) ? __fsub_rn(1.0f, tf) : tf;
                // ---- first-layer operand A1 (over A_h) and the tangent seeds (A_u, A_v chunk 0) ----
                build_first_operand<DOMAIN, TANGENTS, H>(f, row, x0, x1, alpha, tg, av_row);
                if (MODE != kModePdf && t == P.T - 1) {      // last read of the record: hand the slot back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.empty[sl]));
                }
                publish_and_issue<TANGENTS, H>(g, q, 0, NH, tg_mma, b_hid0, b_out, av_desc, bar_d);

                // ---- activation rounds: layer 1 and the hidden layers share one instruction stream ----
#pragma unroll 1
                for (int l = 0; l < NH; ++l) {
                    mbar_wait(bar_d, pd); pd ^= 1u;
                    tc_fence_after();
                    TC_TRACE(0);
                    activation_pass<TANGENTS, ACT, H>(tg, av_row, trace_r, g);
                    publish_and_issue<TANGENTS, H>(g, q, l + 1, NH, tg_mma, b_hid0, b_out, av_desc, bar_d, trace_r);
#ifdef BSDFDIFF_TC_TRACE
                    ++trace_r;
#endif
                }

                // ---- output round: d, dd/dx0, dd/dx1 ----
                mbar_wait(bar_d, pd); pd ^= 1u;
                tc_fence_after();
                float d0, d1, du0 = 0.f, du1 = 0.f, dv0 = 0.f, dv1 = 0.f;
                tmem_ld2(tg + kColDz, d0, d1);
                if (TANGENTS) { tmem_ld2(tg + kColDu, du0, du1); tmem_ld2(tg + kColDv, dv0, dv1); }
                tc_wait_ld();
                if (TANGENTS) {
                    const float j00 = __fmaf_rn(step, du0, 1.0f), j01 = __fmul_rn(step, dv0);
                    const float j10 = __fmul_rn(step, du1), j11 = __fmaf_rn(step, dv1, 1.0f);
                    const float det = __fmaf_rn(j00, j11, -__fmul_rn(j01, j10));
                    R = (MODE == kModePdf) ? __fmul_rn(R, det) : __fdividef(R, det);
                    cond.step(j00, j01, j10, j11, det);
                }
                x0 = fmaf(step, d0, x0);
                x1 = fmaf(step, d1, x1);
            }

            if (MODE == kModeSample) {
                if (valid) store_sample<true>(P, i, x0, x1, p0 * R);
                if (P.fix_thr != 0.0f)
                    flag_for_fixup(P, valid && cond.weight() < (MULTI ? material_fix_thr(P, mat, false) : P.fix_thr), i,
                                   MULTI ? mat : -1, theta_o, x1_start);
            } else if (MODE == kModePdf) {
                float bp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bp[j] = f[kFBp + j][row];
                const float wiz = f[kFWiz][row], wox = f[kFWo][row], woy = f[kFWo + 1][row], woz = f[kFWo + 2][row];
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&S.empty[sl]));
                if (valid)
                    store_pdf<true>(P, i, __fmul_rn(__expf(base_logprob_fast<DOMAIN>(bp, x0, x1)), R), wiz, wox, woy, woz, theta_o);
                if (P.fix_thr != 0.0f) {
                    const float kappa = (DOMAIN == kDisk) ? 0.0f : softplus_fast(bp[3]) + 1e-3f;
                    const float gn = base_grad_norm(DOMAIN, bp, kappa, x0, x1);
                    flag_for_fixup(P, valid && cond.weight() * fminf(1.0f, __fdividef(25.0f, gn)) <
                                          (MULTI ? material_fix_thr(P, mat, true) : P.fix_thr), i,
                                   MULTI ? mat : -1);
                }
            } else {
                if (valid) reinterpret_cast<float2*>(P.out_dir)[i] = make_float2(x0, x1);
            }
            sl += kGroups;
            if (sl >= kSlots) { sl -= kSlots; ++use; }
        }
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

template <int DOMAIN, int MODE, int ACT, int H, bool MULTI = false>
static int launch_tc_t(const FlowParams& P, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // multi-material launches: the number of virtual tiles is only known on the device (<= n / 128 + n_materials)
    const long long tiles = MULTI ? P.n / kTile + P.n_materials : (P.n + kTile - 1) / kTile;
    long long grid = sms;
    if (grid > tiles) grid = tiles;
    if (grid < 1) return 0;
    const size_t smem = tc_smem_bytes(H, P.n_hidden, MULTI);
    if (smem > 227u * 1024u) return -2;
    // the opt-in is per DEVICE (a process may drive several GPUs): set it on every launch, like launch_simt_t does --
    // cheap and CUDA-graph-capture safe
    if (cudaFuncSetAttribute(flow_tc_kernel<DOMAIN, MODE, ACT, H, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) return -3;
    flow_tc_kernel<DOMAIN, MODE, ACT, H, MULTI><<<(unsigned)grid, kTcThreads, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

template <int DOMAIN, int ACT>
static int launch_tc_m(const FlowParams& P, cudaStream_t stream) {
    if (P.n_materials > 0) {                  // one wavefront, several materials: sampler shapes only
        if (P.hidden != 32 || kDuo || kSmemAv) return -2;
#if !BSDFDIFF_TC_DUO && BSDFDIFF_TC_GROUPS != 5
        if (P.mode == kModeSample) return launch_tc_t<DOMAIN, kModeSample, ACT, 32, true>(P, stream);
        if (P.mode == kModePdf) return launch_tc_t<DOMAIN, kModePdf, ACT, 32, true>(P, stream);
#endif
        return -2;
    }
    if (P.hidden == 64) {                     // reflow teacher nets (NN_cond_pos_spherical_complicate): forward only
        if (P.mode != kModeForward) return -2;
        return launch_tc_t<DOMAIN, kModeForward, ACT, 64>(P, stream);
    }
    switch (P.mode) {
        case kModeSample: return launch_tc_t<DOMAIN, kModeSample, ACT, 32>(P, stream);
        case kModePdf: return launch_tc_t<DOMAIN, kModePdf, ACT, 32>(P, stream);
        default: return launch_tc_t<DOMAIN, kModeForward, ACT, 32>(P, stream);
    }
}

// variant 1 = tc16 (tanh.approx activation), 2 = tc16 with fp32 exp-form activation (cross-check)
int launch_tc(const FlowParams& P, cudaStream_t stream, int variant) {
    {   // tcgen05 / TMEM exist on compute capability 10.x only
        int dev = 0, major = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -3;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (major != 10) return -4;
    }
    if ((P.hidden != 32 && P.hidden != 64) || P.n_hidden < 2 || P.n_hidden > 6) return -2;   // else: CUDA-core path
    if (P.domain == kDisk)
        return variant == 2 ? launch_tc_m<kDisk, 0>(P, stream) : launch_tc_m<kDisk, 1>(P, stream);
    return variant == 2 ? launch_tc_m<kSpherical, 0>(P, stream) : launch_tc_m<kSpherical, 1>(P, stream);
}

int tc_trace_read(unsigned long long* out, int max_words) {
#ifdef BSDFDIFF_TC_TRACE
    const int words = kTraceWarps * kTraceRounds * kTraceEvents;
    if (max_words < words) return -1;
    if (cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * words) != cudaSuccess) return -3;
    return words;
#else
    (void)out; (void)max_words;
    return 0;
#endif
}

unsigned int tc_timeout_flag() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_tc_timeout_flag, sizeof(v));
    return v;
}

}  // namespace bsdfdiff
