// Ground-truth evaluator for the RGL measured-BSDF tensor files (rendering/measuredbsdfs/*.bsdf) on the GPU, and the
// plugins' throughput weight + firefly clamp fused on top of it.
//
// The reference's measured plugins call Mitsuba 3's `measured` BSDF for this (self.bsdf.eval(ctx, si, bs.wo),
// rendering/brdf_measured_disk.py:92, rendering/brdf_measured_spherical.py:100) and then round-trip Dr.Jit -> torch ->
// Dr.Jit twice for the firefly clamp (:93-100 / :101-108).  Mitsuba is third-party and not vendored; this file
// implements the published model (Dupuy & Jakob 2018; Mitsuba 3 src/bsdfs/measured.cpp eval() and
// include/mitsuba/core/distr_2d.h Marginal2D<Float, Dimension, Continuous = true>) -- see oracle/measured_oracle.py,
// which restates the same arithmetic in numpy and is what the tests compare this kernel against:
//     eval(wi, wo) = rgb(vndf^-1(u_m) | phi_i, theta_i) * ndf(u_m) / (4 sigma(u_wi)),   m = normalize(wi + wo),
//     u = (sqrt(2 theta / pi), (phi + pi) / 2 pi);  isotropic data: phi_m relative to phi_i.
// One thread per query; the tables (0.66 MB isotropic, 4.5 MB anisotropic) are read-only and L2-resident, so the kernel
// is bound by ~90 dependent 4-byte gathers per query, not by HBM (44 B/query of streaming traffic).
#include "../../include/bsdfdiff.h"
#include "common.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace bsdfdiff {

constexpr uint32_t kMeasuredMagic = 0xB5DF3EA5u;

struct MeasuredHeader {
    uint32_t magic;
    int32_t n_phi, n_theta, isotropic, reduction, jacobian;
    int32_t ndf_w, ndf_h, sig_w, sig_h, v_w, v_h, r_w, r_h;
    uint32_t off_phi, off_theta, off_ndf, off_sigma, off_vdata, off_vcond, off_vmarg, off_rgb;   // in floats from the blob start
    uint32_t total_bytes;
    uint32_t reserved[9];
};
static_assert(sizeof(MeasuredHeader) == 128, "measured header is 128 bytes");

struct ParamCell {           // one axis of the (phi_i, theta_i) parameter grid: cell index and the weight of its upper node
    int i0, i1;
    float w1;
};

// math::find_interval + the clamped linear weight of Marginal2D::eval (distr_2d.h)
__device__ __forceinline__ ParamCell param_cell(const float* __restrict__ v, int n, float x) {
    ParamCell c;
    if (n == 1) { c.i0 = c.i1 = 0; c.w1 = 0.0f; return c; }
    int lo = 0, hi = n - 2;                       // largest i in [0, n-2] with v[i] <= x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(v + mid) <= x) lo = mid; else hi = mid - 1;
    }
    const float p0 = __ldg(v + lo), p1 = __ldg(v + lo + 1);
    c.i0 = lo; c.i1 = lo + 1;
    c.w1 = fminf(fmaxf((x - p0) / (p1 - p0), 0.0f), 1.0f);
    return c;
}

// multilinear interpolation over the parameter axes of one table entry; slice index = iphi * n_theta + itheta
// (times `per` consecutive sub-slices, e.g. the 3 colour channels, offset `sub`)
struct Slices {
    size_t s[4];
    float w[4];
};
__device__ __forceinline__ Slices make_slices(const ParamCell& ph, const ParamCell& th, int n_theta, size_t slice_elems,
                                              int per, int sub) {
    Slices S;
    const int ip[2] = {ph.i0, ph.i1}, it[2] = {th.i0, th.i1};
    const float wp[2] = {1.0f - ph.w1, ph.w1}, wt[2] = {1.0f - th.w1, th.w1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            S.s[2 * a + b] = ((size_t)(ip[a] * n_theta + it[b]) * per + sub) * slice_elems;
            S.w[2 * a + b] = wp[a] * wt[b];
        }
    return S;
}
__device__ __forceinline__ float lookup(const float* __restrict__ t, const Slices& S, size_t idx) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (S.w[k] != 0.0f) v = fmaf(S.w[k], __ldg(t + S.s[k] + idx), v);
    return v;
}

// bilinear interpolation of a [h, w] grid at pos in [0,1]^2 (x along w)
template <bool SLICED>
__device__ __forceinline__ float grid_eval(const float* __restrict__ t, const Slices* S, int w, int h, float px, float py) {
    px *= (float)(w - 1); py *= (float)(h - 1);
    const int ox = min(max((int)px, 0), w - 2), oy = min(max((int)py, 0), h - 2);
    const float wx1 = px - (float)ox, wy1 = py - (float)oy;
    const size_t i = (size_t)oy * w + ox;
    float v00, v10, v01, v11;
    if (SLICED) {
        v00 = lookup(t, *S, i); v10 = lookup(t, *S, i + 1); v01 = lookup(t, *S, i + w); v11 = lookup(t, *S, i + w + 1);
    } else {
        v00 = __ldg(t + i); v10 = __ldg(t + i + 1); v01 = __ldg(t + i + w); v11 = __ldg(t + i + w + 1);
    }
    return fmaf(1.0f - wy1, fmaf(1.0f - wx1, v00, wx1 * v10), wy1 * fmaf(1.0f - wx1, v01, wx1 * v11));
}

__device__ __forceinline__ float elevation(float x, float y, float z) {
    return 2.0f * asinf(fminf(0.5f * sqrtf(x * x + y * y + (z - 1.0f) * (z - 1.0f)), 1.0f));
}

// Mitsuba `measured` eval for one (wi, wo) pair; returns false (and zero) outside cos_i > 0 && cos_o > 0
__device__ __forceinline__ bool measured_eval_one(const unsigned char* __restrict__ blob, float wix, float wiy, float wiz,
                                                  float wox, float woy, float woz, float rgb[3]) {
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    if (!(wiz > 0.0f && woz > 0.0f)) return false;
    const MeasuredHeader* H = reinterpret_cast<const MeasuredHeader*>(blob);
    const float* base = reinterpret_cast<const float*>(blob);
    if (H->reduction >= 2) {                                   // data covers half / a quarter of the azimuth range
        const float sy = wiy, sx = (H->reduction == 4) ? wix : sy;
        if (sx < 0.0f) { wix = -wix; wox = -wox; }
        if (sy < 0.0f) { wiy = -wiy; woy = -woy; }
    }
    float mx = wix + wox, my = wiy + woy, mz = wiz + woz;
    const float inv = rsqrtf(mx * mx + my * my + mz * mz);
    mx *= inv; my *= inv; mz *= inv;
    const float theta_i = elevation(wix, wiy, wiz), phi_i = atan2f(wiy, wix);
    const float theta_m = elevation(mx, my, mz), phi_m = atan2f(my, mx);
    const float kInvPi2 = 0.63661977236758134f, kInv2Pi = 0.15915494309189535f, kPi = 3.14159265358979323846f;
    const float u_wi_x = sqrtf(theta_i * kInvPi2), u_wi_y = (phi_i + kPi) * kInv2Pi;
    const float u_m_x = sqrtf(theta_m * kInvPi2);
    float u_m_y = ((H->isotropic ? (phi_m - phi_i) : phi_m) + kPi) * kInv2Pi;
    u_m_y -= floorf(u_m_y);

    const ParamCell ph = param_cell(base + H->off_phi, H->n_phi, phi_i);
    const ParamCell th = param_cell(base + H->off_theta, H->n_theta, theta_i);

    // ---- sample = vndf.invert(u_m | phi_i, theta_i): u_x = conditional CDF, u_y = marginal CDF of the bilinear pdf ----
    const int vw = H->v_w, vh = H->v_h;
    const Slices Sd = make_slices(ph, th, H->n_theta, (size_t)vw * vh, 1, 0);
    const Slices Sc = make_slices(ph, th, H->n_theta, (size_t)(vw - 1) * vh, 1, 0);
    const Slices Sm = make_slices(ph, th, H->n_theta, (size_t)(vh - 1), 1, 0);
    const float* vdata = base + H->off_vdata;
    const float* vcond = base + H->off_vcond;
    const float* vmarg = base + H->off_vmarg;
    float px = u_m_x * (float)(vw - 1), py = u_m_y * (float)(vh - 1);
    const int ox = min(max((int)px, 0), vw - 2), oy = min(max((int)py, 0), vh - 2);
    const float sx = px - (float)ox, sy = py - (float)oy;
    const size_t i = (size_t)oy * vw + ox;
    const float v00 = lookup(vdata, Sd, i), v10 = lookup(vdata, Sd, i + 1), v01 = lookup(vdata, Sd, i + vw),
                v11 = lookup(vdata, Sd, i + vw + 1);
    const float c0 = fmaf(1.0f - sy, v00, sy * v01), c1 = fmaf(1.0f - sy, v10, sy * v11);
    const float inv_area = 1.0f / ((float)(vw - 1) * (float)(vh - 1));
    const float part_x = (sx * c0 + 0.5f * sx * sx * (c1 - c0)) * inv_area;
    const size_t r0 = (size_t)oy * (vw - 1), r1 = (size_t)(oy + 1) * (vw - 1);
    float left0 = 0.0f, left1 = 0.0f;
    if (ox > 0) { left0 = lookup(vcond, Sc, r0 + ox - 1); left1 = lookup(vcond, Sc, r1 + ox - 1); }
    const float row0 = lookup(vcond, Sc, r0 + vw - 2), row1 = lookup(vcond, Sc, r1 + vw - 2);
    const float smp_x = (fmaf(1.0f - sy, left0, sy * left1) + part_x) / fmaf(1.0f - sy, row0, sy * row1);
    const float below = (oy > 0) ? lookup(vmarg, Sm, (size_t)(oy - 1)) : 0.0f;
    const float smp_y = below + sy * row0 + 0.5f * sy * sy * (row1 - row0);

    // ---- spectra at the warped position, NDF / projected-area Jacobian ----
    float scale = 1.0f;
    if (H->jacobian) {
        const float d = grid_eval<false>(base + H->off_ndf, nullptr, H->ndf_w, H->ndf_h, u_m_x, u_m_y);
        const float s = grid_eval<false>(base + H->off_sigma, nullptr, H->sig_w, H->sig_h, u_wi_x, u_wi_y);
        scale = d / (4.0f * s);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const Slices Sr = make_slices(ph, th, H->n_theta, (size_t)H->r_w * H->r_h, 3, c);
        rgb[c] = grid_eval<true>(base + H->off_rgb, &Sr, H->r_w, H->r_h, smp_x, smp_y) * scale;
    }
    return true;
}

__global__ void __launch_bounds__(256) measured_eval_kernel(const unsigned char* __restrict__ blob, long long n,
                                                            const float* __restrict__ wi, const float* __restrict__ wo,
                                                            float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float rgb[3];
        measured_eval_one(blob, wi[3 * i], wi[3 * i + 1], wi[3 * i + 2], wo[3 * i], wo[3 * i + 1], wo[3 * i + 2], rgb);
        out[3 * i] = rgb[0]; out[3 * i + 1] = rgb[1]; out[3 * i + 2] = rgb[2];
    }
}

// Tail of MyBSDF.sample of the two measured plugins, after the sampler kernel has produced (wo, bs_pdf):
//   value = eval(wi, wo) / bs_pdf * albedo                                   brdf_measured_disk.py:92-93, _spherical.py:100-101
//   spherical only: value = select(cos_i > 0 & bs_pdf > 0, value, 0)         _spherical.py:104
//   pdf   = lum(value) < clamp ? bs_pdf : 0     (firefly clamp; NaN compares false)   :97-100 / :105-108
//   weight = select(cos_i > 0 & pdf > 0 & cos_o > 0, value, 0)              :101 / :109
// The reference clamps the DOMAIN pdf and re-multiplies by cos / (1/sin); bs_pdf already carries that factor, and
// zeroing either is the same.
__global__ void __launch_bounds__(256) measured_weight_kernel(const unsigned char* __restrict__ blob, int kind, long long n,
                                                              const float* __restrict__ wi, const float* __restrict__ wo,
                                                              const float* __restrict__ bs_pdf, float a0, float a1, float a2,
                                                              float clamp, float* __restrict__ out_w,
                                                              float* __restrict__ out_pdf) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float wiz = wi[3 * i + 2], woz = wo[3 * i + 2], p = bs_pdf[i];
        float rgb[3];
        measured_eval_one(blob, wi[3 * i], wi[3 * i + 1], wiz, wo[3 * i], wo[3 * i + 1], woz, rgb);
        float v0 = rgb[0] / p * a0, v1 = rgb[1] / p * a1, v2 = rgb[2] / p * a2;
        if (kind == kEpiSpherical && !(wiz > 0.0f && p > 0.0f)) { v0 = v1 = v2 = 0.0f; }
        const float lum = 0.2126f * v0 + 0.7152f * v1 + 0.0722f * v2;            // rgb2lum, utils/mitsuba_brdf_draw.py:36-38
        const float pdf = (lum < clamp) ? p : 0.0f;
        const bool keep = (wiz > 0.0f) && (pdf > 0.0f) && (woz > 0.0f);
        out_w[3 * i] = keep ? v0 : 0.0f; out_w[3 * i + 1] = keep ? v1 : 0.0f; out_w[3 * i + 2] = keep ? v2 : 0.0f;
        out_pdf[i] = pdf;
    }
}

static int grid_for(long long n) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (n + 255) / 256, cap = (long long)sms * 8;
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace bsdfdiff

using namespace bsdfdiff;

static inline size_t up32(size_t floats) { return (floats + 31) / 32 * 32; }     // keep every table 128-byte aligned

extern "C" size_t bsdfdiff_measured_blob_bytes(int n_phi, int n_theta, int ndf_w, int ndf_h, int sig_w, int sig_h,
                                                int v_w, int v_h, int r_w, int r_h) {
    if (n_phi < 1 || n_theta < 1 || ndf_w < 2 || ndf_h < 2 || sig_w < 2 || sig_h < 2 || v_w < 2 || v_h < 2 || r_w < 2 || r_h < 2)
        return 0;
    const size_t S = (size_t)n_phi * n_theta;
    size_t f = sizeof(MeasuredHeader) / 4;
    f += up32(n_phi) + up32(n_theta) + up32((size_t)ndf_w * ndf_h) + up32((size_t)sig_w * sig_h);
    f += up32(S * v_w * v_h) + up32(S * (v_w - 1) * v_h) + up32(S * (v_h - 1)) + up32(S * 3 * r_w * r_h);
    return f * 4;
}

// Host-side packing: raw tensor-file fields -> device blob.  The VNDF is normalised and its conditional / marginal CDFs are
// built exactly as Marginal2D's constructor does (trapezoid sums accumulated in double, stored as float).
extern "C" int bsdfdiff_measured_pack(const float* phi_i, int n_phi, const float* theta_i, int n_theta, const float* ndf,
                                      int ndf_w, int ndf_h, const float* sigma, int sig_w, int sig_h, const float* vndf,
                                      int v_w, int v_h, const float* rgb, int r_w, int r_h, int jacobian, void* blob_out) {
    const size_t total = bsdfdiff_measured_blob_bytes(n_phi, n_theta, ndf_w, ndf_h, sig_w, sig_h, v_w, v_h, r_w, r_h);
    if (!total || !phi_i || !theta_i || !ndf || !sigma || !vndf || !rgb || !blob_out) return BSDFDIFF_EINVAL;
    const size_t S = (size_t)n_phi * n_theta;
    std::memset(blob_out, 0, total);
    MeasuredHeader H{};
    H.magic = kMeasuredMagic; H.n_phi = n_phi; H.n_theta = n_theta; H.isotropic = n_phi <= 2; H.jacobian = jacobian != 0;
    H.reduction = H.isotropic ? 0 : (int)std::lrint(2.0 * 3.14159265358979323846 / ((double)phi_i[n_phi - 1] - (double)phi_i[0]));
    H.ndf_w = ndf_w; H.ndf_h = ndf_h; H.sig_w = sig_w; H.sig_h = sig_h; H.v_w = v_w; H.v_h = v_h; H.r_w = r_w; H.r_h = r_h;
    size_t f = sizeof(MeasuredHeader) / 4;
    H.off_phi = (uint32_t)f; f += up32(n_phi);
    H.off_theta = (uint32_t)f; f += up32(n_theta);
    H.off_ndf = (uint32_t)f; f += up32((size_t)ndf_w * ndf_h);
    H.off_sigma = (uint32_t)f; f += up32((size_t)sig_w * sig_h);
    H.off_vdata = (uint32_t)f; f += up32(S * v_w * v_h);
    H.off_vcond = (uint32_t)f; f += up32(S * (v_w - 1) * v_h);
    H.off_vmarg = (uint32_t)f; f += up32(S * (v_h - 1));
    H.off_rgb = (uint32_t)f; f += up32(S * 3 * r_w * r_h);
    H.total_bytes = (uint32_t)total;
    float* out = static_cast<float*>(blob_out);
    std::memcpy(out, &H, sizeof(H));
    std::memcpy(out + H.off_phi, phi_i, sizeof(float) * n_phi);
    std::memcpy(out + H.off_theta, theta_i, sizeof(float) * n_theta);
    std::memcpy(out + H.off_ndf, ndf, sizeof(float) * ndf_w * ndf_h);
    std::memcpy(out + H.off_sigma, sigma, sizeof(float) * sig_w * sig_h);
    std::memcpy(out + H.off_rgb, rgb, sizeof(float) * S * 3 * r_w * r_h);
    std::vector<double> cond((size_t)(v_w - 1) * v_h), marg(v_h - 1);
    for (size_t s = 0; s < S; ++s) {
        const float* d = vndf + s * v_w * v_h;
        for (int y = 0; y < v_h; ++y) {
            double sum = 0.0;
            for (int x = 0; x < v_w - 1; ++x) {
                sum += 0.5 * ((double)d[(size_t)y * v_w + x] + (double)d[(size_t)y * v_w + x + 1]);
                cond[(size_t)y * (v_w - 1) + x] = sum;
            }
        }
        double sum = 0.0;
        for (int y = 0; y < v_h - 1; ++y) {
            sum += 0.5 * (cond[(size_t)(y + 1) * (v_w - 1) - 1] + cond[(size_t)(y + 2) * (v_w - 1) - 1]);
            marg[y] = sum;
        }
        const double norm = 1.0 / marg[v_h - 2];
        float* od = out + H.off_vdata + s * v_w * v_h;
        float* oc = out + H.off_vcond + s * (size_t)(v_w - 1) * v_h;
        float* om = out + H.off_vmarg + s * (size_t)(v_h - 1);
        const double dn = norm * (double)(v_w - 1) * (double)(v_h - 1);
        for (size_t k = 0; k < (size_t)v_w * v_h; ++k) od[k] = (float)((double)d[k] * dn);
        for (size_t k = 0; k < cond.size(); ++k) oc[k] = (float)(cond[k] * norm);
        for (size_t k = 0; k < marg.size(); ++k) om[k] = (float)(marg[k] * norm);
    }
    return BSDFDIFF_OK;
}

extern "C" int bsdfdiff_measured_eval(const void* blob, int64_t n, const float* wi, const float* wo, float* out_rgb,
                                      void* cuda_stream) {
    if (n < 0 || !blob) return BSDFDIFF_EINVAL;
    if (n == 0) return BSDFDIFF_OK;
    if (!wi || !wo || !out_rgb) return BSDFDIFF_EINVAL;
    measured_eval_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const unsigned char*>(blob), n, wi, wo, out_rgb);
    return cudaGetLastError() == cudaSuccess ? BSDFDIFF_OK : BSDFDIFF_ECUDA;
}

extern "C" int bsdfdiff_measured_weight(const void* blob, int epilogue, int64_t n, const float* wi, const float* wo,
                                        const float* bs_pdf, float albedo_r, float albedo_g, float albedo_b, float clamp,
                                        float* out_weight, float* out_pdf, void* cuda_stream) {
    if (n < 0 || !blob || (epilogue != kEpiDisk && epilogue != kEpiSpherical)) return BSDFDIFF_EINVAL;
    if (n == 0) return BSDFDIFF_OK;
    if (!wi || !wo || !bs_pdf || !out_weight || !out_pdf) return BSDFDIFF_EINVAL;
    measured_weight_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const unsigned char*>(blob), epilogue, n, wi, wo, bs_pdf, albedo_r, albedo_g, albedo_b, clamp, out_weight,
        out_pdf);
    return cudaGetLastError() == cudaSuccess ? BSDFDIFF_OK : BSDFDIFF_ECUDA;
}
