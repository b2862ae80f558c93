// Device-side plan of a multi-material wavefront (multi.cu): layout view + launchers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace bsdfdiff {

constexpr int kMaxMaterials = 255;
constexpr int kPlanHeaderWords = 1280;

struct MultiPlanView {
    unsigned int *counts, *cursor, *seg_off, *n_tiles, *fix_count;
    int4* tiles;
    unsigned int *perm, *fix_list;
    float* x0;
};

size_t multi_scratch_bytes(long long n, int n_materials);
MultiPlanView multi_view(void* scratch, long long n, int n_materials);
int launch_multi_plan(long long n, const int* material_id, int n_materials, void* scratch, cudaStream_t stream);
int launch_multi_zero_inactive(long long n, int n_materials, void* scratch, float* out_dir, int dir_cols, float* out_pdf,
                               cudaStream_t stream);

}  // namespace bsdfdiff
