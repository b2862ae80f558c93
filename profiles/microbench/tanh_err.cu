// Error profile of tanh.approx.f32 on sm_100a against double tanh: relative error of t and of (1 - |t|) (the factor the
// silu derivative multiplies by h) per range of |x|.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tanh_err tanh_err.cu
#include <cmath>
#include <cstdio>
#include <vector>
__global__ void k(const float* x, float* t, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x[i])); t[i] = r; }
}
int main() {
    const int n = 1 << 22;
    std::vector<float> hx(n), ht(n);
    for (int i = 0; i < n; ++i) hx[i] = 12.0f * (float)i / n;
    float *dx, *dt;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dt, n * 4);
    cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dt, n);
    cudaMemcpy(ht.data(), dt, n * 4, cudaMemcpyDeviceToHost);
    const double edges[] = {0, 0.125, 0.25, 0.5, 1, 1.5, 2, 2.5, 3, 4, 5, 6, 8, 12};
    printf("%-14s %12s %12s %14s %14s\n", "|x| range", "max rel(t)", "rms rel(t)", "max rel(1-t)", "rms rel(1-t)");
    for (int b = 0; b + 1 < (int)(sizeof(edges) / sizeof(double)); ++b) {
        double mr = 0, sr = 0, m1 = 0, s1 = 0; long c = 0;
        for (int i = 1; i < n; ++i) {
            if (hx[i] < edges[b] || hx[i] >= edges[b + 1]) continue;
            const double te = tanh((double)hx[i]), ta = ht[i];
            const double r = fabs(ta - te) / te, r1 = fabs((1.0 - ta) - (1.0 - te)) / (1.0 - te);
            mr = fmax(mr, r); sr += r * r; m1 = fmax(m1, r1); s1 += r1 * r1; ++c;
        }
        printf("[%5.3f,%6.3f) %12.3e %12.3e %14.3e %14.3e\n", edges[b], edges[b + 1], mr, sqrt(sr / c), m1, sqrt(s1 / c));
    }
    return 0;
}
