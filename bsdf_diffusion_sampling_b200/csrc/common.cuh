// Shared device/host definitions for libbsdfdiff: packed-weight blob layout, launch parameters,
// Philox4x32-10, base-distribution net, domain mappings and plugin epilogues.
//
// Arithmetic spec: SURVEY.md Appendix A (validated against the reference's autograd path).
// Reference lines are cited next to each function.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <float.h>

namespace bsdfdiff {

constexpr int kDisk = 0, kSpherical = 1;
constexpr int kEpiRaw = 0, kEpiDisk = 1, kEpiSpherical = 2, kEpiBsdf = 3;
constexpr int kModeSample = 0, kModePdf = 1, kModeForward = 2;
constexpr int kBaseFloats = 308;
constexpr int kPE5 = 22;           // 2 + 4*5
constexpr int kPE3 = 14;           // 2 + 4*3
constexpr uint32_t kMagic = 0xB5DFD1F0u;

// ---------------------------------------------------------------------------------------------
// Packed flow-net blob (built on the host by bsdfdiff_pack_flow, resident in HBM, staged into
// shared memory once per persistent CTA).
//   [header 64 B]
//   [fp32 image]  layer-major, each layer TRANSPOSED to [k][j] (input-major) so that a warp reads
//                 consecutive output neurons j as float4 broadcasts:
//                 W1t [in_dim][H], W2t.. [H][H], Woutt [H][2]
//   [fp16 image]  tensor-core B operands, one per layer, in the UMMA canonical K-major no-swizzle
//                 core-matrix layout (8 rows x 16 B per core matrix):
//                 byte(n,k) = ((k/8)*(N/8) + n/8)*128 + (n%8)*16 + (k%8)*2
//                 first layer K padded to 32 with hi/lo split columns (see flow_tc.cu), output layer
//                 N padded to 16.  All TC weights pre-scaled by 0.5 except the output layer
//                 (tanh-form sigmoid: the MMA produces z/2 directly; see flow_tc.cu).
// ---------------------------------------------------------------------------------------------
struct PackedHeader {
    uint32_t magic;
    int32_t in_dim;       // 25 | 26 (or arbitrary <= 32 for tcnn-style nets)
    int32_t hidden;       // 32 | 64
    int32_t n_hidden;     // number of hidden layers (3 disk, 4 spherical, 6 reflow-complex)
    int32_t domain;       // kDisk | kSpherical
    uint32_t off_f32;     // byte offset of the fp32 image
    uint32_t f32_bytes;
    uint32_t off_f16;     // byte offset of the fp16 image
    uint32_t f16_bytes;
    uint32_t total_bytes;
    // reserved[0..2]: optional per-material fix-up thresholds (host: weights.PackedFlow.set_fixup): kFixThrMagic, then the
    // conditioning thresholds of sample() and pdf() as float bits.  Read by the multi-material launches when the caller
    // passes a NEGATIVE fix_threshold ("every material's own threshold; |fix_threshold| where a blob carries none").
    uint32_t reserved[6];
};
constexpr uint32_t kFixThrMagic = 0x54584946u;   // "FIXT"
static_assert(sizeof(PackedHeader) == 64, "header must be 64 bytes");

__host__ __device__ inline int f32_image_floats(int in_dim, int H, int n_hidden) {
    return in_dim * H + (n_hidden - 1) * H * H + 2 * H;
}
// fp16 image: per layer a HI operand followed by a LO operand (W = hi + lo, each fp16; the value path
// multiplies by both so weights carry ~22 mantissa bits, the tangent path uses hi only):
// first layer 2 x [H x 32], hidden 2 x [H x H] x (n_hidden-1), output 2 x [16 x H]
__host__ __device__ inline int f16_image_halves(int H, int n_hidden) {
    return 2 * (H * 32 + (n_hidden - 1) * H * H + 16 * H);
}

// Addressing of a 3-component direction tensor (plugin epilogues): component c of row i lives at
// base[i * rs + (c == 0 ? 0 : c == 1 ? c1 : c2)] (offsets in floats).  Interleaved [n,3]: {3, 1, 2}.  Planar / three
// separate arrays (Dr.Jit's Vector3f is three arrays: zero-copy through DLPack): {1, y - x, z - x}.
struct Dir3Layout {
    long long rs, c1, c2;
};
struct FlowParams {
    int domain, mode, epilogue, T;
    int in_dim, hidden, n_hidden;
    long long n;
    long long wi_repeat;          // query i reads wi[i / wi_repeat]
    const float* wi;              // [.,2] or [.,3]
    const float* wo;              // pdf mode
    Dir3Layout wi_l, wo_l, out_l; // layouts of wi / wo / out_dir for the 3-component (plugin) epilogues
    const float* x0;              // replayed base sample / reflow start (may be null)
    const float* u_noise;         // [n,3] uniforms in [0,1) that replace the Philox draws (Mitsuba's sample2.x, sample2.y, sample1;
                                  // brdf_measured_disk.py:59 hands them to sample() and the reference ignores them), or null
    const unsigned char* flow;    // packed blob (device)
    const float* base;            // 308 floats (device) or null (forward mode with x0 given)
    unsigned long long seed, offset;
    long long first_index;
    float* out_dir;               // [n,2] or [n,3]
    float* out_pdf;               // [n]
    float* out_x0;                // optional [n,2]
    // conditioning-triggered fp32 fix-up (PREC_TC16 only): the tensor-core kernel appends the row index of
    // every query whose pdf is ill-conditioned w.r.t. fp16 rounding (weight < fix_thr, see cond_weight()) to
    // fix_list[atomicAdd(fix_count)]; the CUDA-core kernel then recomputes exactly those rows in fp32
    float fix_thr;                // 0 = off; < 0 (multi-material launches): each material's own threshold from its blob header,
                                  // |fix_thr| where the blob carries none (material_fix_thr)
    unsigned int* fix_count;      // device counter (zeroed by the caller before the tensor-core launch)
    unsigned int* fix_list;       // device list, capacity n
    float* fix_x0;                // [n,2] base samples of the flagged rows, parallel to fix_list (sample mode): the tensor-core
                                  // kernel stores them next to the row index, the fix-up pass replays them -- no [n,2]
                                  // side buffer of ALL base samples is written just to recompute a few rows
    int fix_pass;                 // 1: this launch IS the fix-up pass (flow_simt_kernel walks fix_list)
    int log_output;               // pdf mode, T == 0, raw epilogue: store log p_base instead of p_base (D_base.log_prob)
    // one wavefront, several materials: rows are bucketed by material ON THE DEVICE (multi.cu) into
    // "virtual tiles" of <= 128 rows of one material; the kernels walk the tile table and switch weight sets per tile
    int n_materials;              // 0 = single material (flow / base above)
    const unsigned char* const* flows;   // device array [n_materials] of packed flow blobs (all of one shape)
    const float* const* bases;           // device array [n_materials] of base-net blobs
    const unsigned int* perm;            // [n] wavefront row of every position of the material-sorted order
    const int4* tiles;                   // per virtual tile {material, first position, rows, first position of the material}
    const unsigned int* n_tiles_dev;     // number of virtual tiles (device scalar)
    const unsigned int* seg_off;         // [n_materials + 1] first position of every material in the sorted order
};

// Conditioning threshold of tile material `mat` in a multi-material launch: P.fix_thr if positive, else the material's own
// (PackedHeader::reserved[1 + pdf]) with |P.fix_thr| as the fallback.
__device__ __forceinline__ float material_fix_thr(const FlowParams& P, int mat, bool pdf_mode) {
    if (P.fix_thr >= 0.0f) return P.fix_thr;
    const PackedHeader* h = reinterpret_cast<const PackedHeader*>(P.flows[mat]);
    return (h->reserved[0] == kFixThrMagic) ? __uint_as_float(h->reserved[pdf_mode ? 2 : 1]) : -P.fix_thr;
}


// Conditioning weight in (0,1] of a query's pdf w.r.t. rounding inside the flow -- the same three factors the
// oracle reports (oracle/bsdf_oracle.c euler(), bsdf_oracle_pdf()):  pdf = p0 (/|*) prod det_t, so an absolute error
// in a step determinant is a relative pdf error ~ 1/|det_t| (factor min(1, min|det| / 0.2)); a state error made early
// is amplified by the later steps' Jacobians (factor min(1, 16 / prod max(1, sigma_max(J_t)))); pdf() evaluates the
// base density at the END of the reverse flow (factor min(1, 25 / |grad log p_base|)).
struct CondTrack {
    float mind, ampsq;            // min_t |det_t|, prod_t max(1, ~sigma_max(J_t)^2)
    __device__ __forceinline__ void reset() { mind = FLT_MAX; ampsq = 1.0f; }
    // sigma_max^2 = (F + sqrt(F^2 - 4 det^2)) / 2 with F = |J|_F^2 = sigma_1^2 + sigma_2^2.  The kernel tracks the
    // square-root-free proxy F - 1 (exact when the smaller singular value is 1, i.e. for the near-identity step maps
    // I + dD/dx / T; a little eager when both exceed 1): 6 FMA-pipe instructions per step and no MUFU op.
    __device__ __forceinline__ void step(float j00, float j01, float j10, float j11, float det) {
        mind = fminf(mind, fabsf(det));
        const float fro = j00 * j00 + j01 * j01 + j10 * j10 + j11 * j11;
        ampsq *= fmaxf(1.0f, fro - 1.0f);
    }
    __device__ __forceinline__ float weight() const {
        return fminf(1.0f, mind * 5.0f) * fminf(1.0f, 16.0f * rsqrtf(ampsq));
    }
};
// |grad_x log p_base(x)| (model.py:393-398 disk, :308-317 spherical)
__device__ __forceinline__ float base_grad_norm(int domain, const float p[4], float kappa, float x0, float x1) {
    float g0, g1;
    if (domain == kDisk) {
        g0 = (x0 - p[0]) * __expf(-2.0f * p[2]);
        g1 = (x1 - p[1]) * __expf(-2.0f * p[3]);
    } else {
        const float sc = __expf(p[1]) + 1e-3f;
        g0 = (x0 - p[0]) / (sc * sc);
        g1 = kappa * __sinf(x1 - p[2]);
    }
    return sqrtf(g0 * g0 + g1 * g1);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (index lo, index hi, offset lo + round, offset hi)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ uint4 philox_draw(unsigned long long seed, unsigned long long offset,
                                             long long index, uint32_t round) {
    unsigned long long off = offset + round;
    uint4 c = make_uint4((uint32_t)index, (uint32_t)((unsigned long long)index >> 32),
                         (uint32_t)off, (uint32_t)(off >> 32));
    return philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// uniform in (0,1), 24 bits
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float r = sqrtf(-2.0f * logf(u01(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    n0 = r * c; n1 = r * s;
}

// ---------------------------------------------------------------------------------------------
// math helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_precise(float z) { return 1.0f / (1.0f + expf(-z)); }
// sigmoid to ~2 ulp without the IEEE division and expf's range reduction: ex2.approx (2 ulp) of a clamped argument, rcp.approx
// plus one Newton step -- 7 instructions instead of ~25.  The fp32 kernels whose instruction count is dominated by the
// activation math use it (flow_lane8.cu: +12 %; train.cu evaluates silu five times per neuron and row).
__device__ __forceinline__ float sigmoid_newton(float z) {
    float e, r;
    const float a = fminf(z * -1.4426950408889634f, 126.0f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
    const float d = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r * fmaf(-d, r, 2.0f);
}

// [v, sin(2^k v), cos(2^k v)]_k, frequency-major, sin before cos, then component (model.py:26,48-57)
template <int L>
__device__ __forceinline__ void positional_encoding(float v0, float v1, float* out) {
    out[0] = v0; out[1] = v1;
    float f = 1.0f;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        float s0, c0, s1, c1;
        sincosf(v0 * f, &s0, &c0);
        sincosf(v1 * f, &s1, &c1);
        out[2 + 4 * k + 0] = s0; out[2 + 4 * k + 1] = s1;
        out[2 + 4 * k + 2] = c0; out[2 + 4 * k + 3] = c1;
        f *= 2.0f;
    }
}

// base net p = Wo silu(W1 PE3(wi) + b1) + bo  (rendering/utils/model.py:382-386); blob in smem/global
__device__ __forceinline__ void base_eval(const float* __restrict__ b, float w0, float w1, float p[4]) {
    float e[kPE3];
    positional_encoding<3>(w0, w1, e);
    p[0] = b[304]; p[1] = b[305]; p[2] = b[306]; p[3] = b[307];
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        float z = b[224 + j];
#pragma unroll
        for (int k = 0; k < kPE3; ++k) z = fmaf(e[k], b[j * kPE3 + k], z);
        float h = z * sigmoid_precise(z);
        p[0] = fmaf(h, b[240 + j], p[0]);
        p[1] = fmaf(h, b[256 + j], p[1]);
        p[2] = fmaf(h, b[272 + j], p[2]);
        p[3] = fmaf(h, b[288 + j], p[3]);
    }
}

__device__ __forceinline__ float softplus_torch(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// torch/distributions/von_mises.py _log_modified_bessel_fn(order=0)
__device__ __forceinline__ float log_i0(float x) {
    if (x < 3.75f) {
        float y = x / 3.75f; y = y * y;
        float r = 0.45813e-2f;
        r = fmaf(y, r, 0.360768e-1f); r = fmaf(y, r, 0.2659732f); r = fmaf(y, r, 1.2067492f);
        r = fmaf(y, r, 3.0899424f);   r = fmaf(y, r, 3.5156229f); r = fmaf(y, r, 1.0f);
        return logf(r);
    }
    float y = 3.75f / x;
    float r = 0.392377e-2f;
    r = fmaf(y, r, -0.1647633e-1f); r = fmaf(y, r, 0.2635537e-1f); r = fmaf(y, r, -0.2057706e-1f);
    r = fmaf(y, r, 0.916281e-2f);   r = fmaf(y, r, -0.157565e-2f); r = fmaf(y, r, 0.225319e-2f);
    r = fmaf(y, r, 0.1328592e-1f);  r = fmaf(y, r, 0.39894228f);
    return x - 0.5f * logf(x) + logf(r);
}

constexpr float kLog2Pi = 1.8378770664093453f;

// log-density of the base distribution at x given the 4 base-net outputs.
//   disk:      model.py:393-398   (diagonal Gaussian, loc=p[0:2], log_scale=p[2:4])
//   spherical: model.py:293-298,308-317 (Gaussian(theta) x vonMises(phi); normaliser uses log_scale as-is)
__device__ __forceinline__ float base_logprob(int domain, const float p[4], float x0, float x1) {
    if (domain == kDisk) {
        float e0 = (x0 - p[0]) / expf(p[2]), e1 = (x1 - p[1]) / expf(p[3]);
        return -kLog2Pi - (p[2] + p[3]) - 0.5f * (e0 * e0 + e1 * e1);
    }
    float kappa = softplus_torch(p[3]) + 1e-3f;
    float e = (x0 - p[0]) / (expf(p[1]) + 1e-3f);
    float loggau = -0.5f * kLog2Pi - p[1] - 0.5f * e * e;
    float logvon = kappa * cosf(x1 - p[2]) - kLog2Pi - log_i0(kappa);
    return loggau + logvon;
}

// Draw the base sample with Philox.
//   disk:      x = loc + eps * exp(log_scale)                         (model.py:387-392)
//   spherical: theta = loc + eps*(exp(ls)+1e-3); phi ~ vonMises(mu,kappa) by Best & Fisher (1979)
//              rejection, wrapped to [-pi,pi)                          (model.py:299-307, torch von_mises.py)
//   The proposal parameter r is evaluated with a cancellation-free rearrangement so fp32 suffices
//   (torch switches to fp64 because its formula cancels for small kappa).
// Renderer-supplied uniforms as the noise source: (ua, ub) feed the Box-Muller pair directly, and the von Mises
// rejection rounds (which need an unbounded number of uniforms) draw from Philox keyed by the BITS of the three
// uniforms -- the base sample is a pure function of the renderer's sample stream.
__device__ __forceinline__ float clamp_u01(float u) { return fminf(fmaxf(u, 2.9802322e-8f), 0.99999994f); }
__device__ __forceinline__ void noise_key(const float* u, unsigned long long& seed, unsigned long long& offset) {
    seed = (unsigned long long)__float_as_uint(u[0]) | ((unsigned long long)__float_as_uint(u[1]) << 32);
    offset = (unsigned long long)__float_as_uint(u[2]) << 8;
}
__device__ __forceinline__ void base_draw(int domain, const float p[4], unsigned long long seed,
                                          unsigned long long offset, long long index, float& x0, float& x1,
                                          const float* u = nullptr) {
    float n0, n1;
    if (u) {
        const float r = sqrtf(-2.0f * logf(clamp_u01(u[0])));
        float s, c;
        sincospif(2.0f * clamp_u01(u[1]), &s, &c);
        n0 = r * c; n1 = r * s;
        noise_key(u, seed, offset);
        index = 0;
    } else {
        uint4 r4 = philox_draw(seed, offset, index, 0);
        box_muller(r4.x, r4.y, n0, n1);
    }
    if (domain == kDisk) {
        x0 = fmaf(n0, expf(p[2]), p[0]);
        x1 = fmaf(n1, expf(p[3]), p[1]);
        return;
    }
    x0 = fmaf(n0, expf(p[1]) + 1e-3f, p[0]);
    const float kappa = softplus_torch(p[3]) + 1e-3f, mu = p[2];
    const float s = sqrtf(fmaf(4.0f * kappa, kappa, 1.0f));
    const float tau = 1.0f + s;
    const float rho = (tau * 2.0f * kappa) / ((s + 1.0f) * (tau + sqrtf(2.0f * tau)));
    const float r = (1.0f + rho * rho) / (2.0f * rho);
    float phi = 0.0f;
    for (uint32_t round = 1; round <= 64; ++round) {                // acceptance >= ~0.66 per round
        uint4 q = philox_draw(seed, offset, index, round);
        float z = cospif(u01(q.x));
        float f = (1.0f + r * z) / (r + z);
        float cc = kappa * (r - f);
        float u2 = u01(q.y);
        phi = ((q.z & 0x80000000u) ? 1.0f : -1.0f) * acosf(fminf(fmaxf(f, -1.0f), 1.0f));
        if ((cc * (2.0f - cc) - u2 > 0.0f) || (logf(cc / u2) + 1.0f - cc >= 0.0f)) break;
    }
    float y = phi + 3.14159265358979f + mu;
    y = y - 6.28318530717959f * floorf(y * 0.159154943091895f);   // python-style modulo
    x1 = y - 3.14159265358979f;
}

// rendering/brdf_measured_spherical.py:35-39
__device__ __forceinline__ void cart_to_spher(float x, float y, float z, float& theta, float& phi) {
    float r = sqrtf(x * x + y * y + z * z);
    theta = acosf(z / (r + 1e-8f));
    phi = atan2f(y, x);
}

// dr.clamp(1/sin_theta(wo), 1, FLT_MAX)  (brdf_measured_spherical.py:89-91; bsdf_myresult.py:81)
__device__ __forceinline__ float inv_sin_clamped(float wx, float wy) {
    float s = sqrtf(fmaxf(wx * wx + wy * wy, 0.0f));
    return fminf(fmaxf(1.0f / s, 1.0f), FLT_MAX);
}

// Load one query's conditioning (domain coords) from the caller's wi tensor.
__device__ __forceinline__ void load_wi(const FlowParams& P, long long i, float& w0, float& w1, float& wz) {
    long long q = (P.wi_repeat > 1) ? i / P.wi_repeat : i;
    if (P.epilogue == kEpiRaw) {
        float2 w = reinterpret_cast<const float2*>(P.wi)[q];
        w0 = w.x; w1 = w.y; wz = 1.0f;
    } else {
        const float* w = P.wi + q * P.wi_l.rs;
        float a = w[0], b = w[P.wi_l.c1], c = w[P.wi_l.c2];
        wz = c;
        if (P.epilogue == kEpiDisk) { w0 = a; w1 = b; }            // brdf_measured_disk.py:66-67
        else cart_to_spher(a, b, c, w0, w1);                        // brdf_measured_spherical.py:76-77
    }
}

// Start state of the reverse (pdf) flow from the caller's wo tensor; woz/sin flag for the masks.
__device__ __forceinline__ void load_wo(const FlowParams& P, long long i, float& x0, float& x1,
                                        float& wox, float& woy, float& woz) {
    if (P.epilogue == kEpiRaw) {
        float2 w = reinterpret_cast<const float2*>(P.wo)[i];
        x0 = w.x; x1 = w.y; wox = woy = 0.0f; woz = 1.0f;
    } else {
        const float* w = P.wo + i * P.wo_l.rs;
        wox = w[0]; woy = w[P.wo_l.c1]; woz = w[P.wo_l.c2];
        if (P.epilogue == kEpiDisk) { x0 = wox; x1 = woy; }        // brdf_measured_disk.py:118-120
        else cart_to_spher(wox, woy, woz, x0, x1);                  // brdf_measured_spherical.py:131-133
    }
}

// Final mapping for sample(): domain state -> outgoing direction + solid-angle pdf (tensor part of
// MyBSDF.sample before the ground-truth firefly clamp).
template <bool FAST = false>   // FAST: MUFU sin/cos (tensor-core path; ~1e-6 abs on wo)
__device__ __forceinline__ void store_sample(const FlowParams& P, long long i, float x0, float x1, float pdf) {
    if (P.epilogue == kEpiRaw) {
        reinterpret_cast<float2*>(P.out_dir)[i] = make_float2(x0, x1);
        P.out_pdf[i] = pdf;
        return;
    }
    float ox, oy, oz;
    if (P.epilogue == kEpiDisk) {                                   // brdf_measured_disk.py:69-82
        bool valid = (x0 * x0 + x1 * x1) < 0.995f;
        if (!valid) { x0 = 0.0f; x1 = 0.0f; pdf = 0.0f; }
        oz = sqrtf(fmaxf(1.0f - (x0 * x0 + x1 * x1), 0.0f));        // disk_to_cart, mitsuba_brdf_draw.py:40-43
        ox = x0; oy = x1;
        pdf = pdf * oz;
    } else {                                                        // brdf_measured_spherical.py:79-92
        float st, ct, sp, cp;
        if (FAST) { __sincosf(x0, &st, &ct); __sincosf(x1, &sp, &cp); }
        else { sincosf(x0, &st, &ct); sincosf(x1, &sp, &cp); }
        if (!(st > 0.00005f)) pdf = 0.0f;
        if (P.epilogue == kEpiSpherical && !(ct > 0.0f)) pdf = 0.0f;
        ox = cp * st; oy = sp * st; oz = ct;                        // sph_to_dir :31-34
        pdf = pdf * inv_sin_clamped(ox, oy);                        // bsdf_myresult.py:81 (abs is a no-op on a sqrt)
    }
    float* o = P.out_dir + i * P.out_l.rs;
    o[0] = ox; o[P.out_l.c1] = oy; o[P.out_l.c2] = oz;
    P.out_pdf[i] = pdf;
}

// Final masks / Jacobian for pdf() (tensor part of MyBSDF.pdf).
template <bool FAST = false>
__device__ __forceinline__ void store_pdf(const FlowParams& P, long long i, float pdf, float wiz,
                                          float wox, float woy, float woz, float theta_o) {
    if (P.epilogue == kEpiDisk) {                                   // brdf_measured_disk.py:122-124
        pdf = (wiz > 0.0f && woz > 0.0f) ? pdf * woz : 0.0f;
    } else if (P.epilogue == kEpiSpherical) {                       // brdf_measured_spherical.py:134-137
        if (!((FAST ? __sinf(theta_o) : sinf(theta_o)) > 0.00005f)) pdf = 0.0f;
        pdf = (wiz > 0.0f && woz > 0.0f) ? pdf * inv_sin_clamped(wox, woy) : 0.0f;
    } else if (P.epilogue == kEpiBsdf) {                            // bsdf_myresult.py:121-130
        pdf = pdf * inv_sin_clamped(wox, woy);
    }
    P.out_pdf[i] = pdf;
}

int launch_simt(const FlowParams& P, cudaStream_t stream);
int launch_lane8(const FlowParams& P, cudaStream_t stream);      // eight lanes per query: 32-wide sampler nets, sample / pdf, T >= 1
int launch_tc(const FlowParams& P, cudaStream_t stream, int variant);
unsigned int tc_timeout_flag();
int tc_trace_read(unsigned long long* out, int max_words);
int launch_mlp_forward_simt(long long n, const float* in, int in_dim, const unsigned char* flow, int H, int n_hidden,
                            float* out, cudaStream_t stream);

}  // namespace bsdfdiff
