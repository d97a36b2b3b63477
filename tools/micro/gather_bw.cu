// Random 8-byte gathers out of an L2-resident table: the access pattern of the hash grid's fine levels (hashencoder.cu:104:
// 8 corners x 2 floats per level and point, table 6 098 108 x 2 fp32 = 46.5 MB).  One thread = one (point, level): eight
// independent gathers at computed (LCG) indices, one 8-byte result.  Prints the sustained gather rate; the roofline the
// hash-grid forward is read against (profiles/README.md).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw.bin gather_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_gather(const float2* __restrict__ table, uint32_t n_entries, float2* __restrict__ out, int64_t n_threads) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_threads) return;
    uint32_t s = (uint32_t)i * 2654435761u + 12345u;
    uint32_t idx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; idx[k] = (uint32_t)(((uint64_t)s * n_entries) >> 32); }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float2 v = __ldg(table + idx[k]); acc.x += v.x; acc.y += v.y; }
    out[i] = acc;
}

int main() {
    const uint32_t entries[2] = {6098108u, 524288u};          // whole table (46.5 MB); one hashed level (4 MB)
    const int64_t n_threads = 262144ll * 16;
    float2 *table, *out;
    cudaMalloc(&table, (size_t)entries[0] * sizeof(float2));
    cudaMalloc(&out, (size_t)n_threads * sizeof(float2));
    cudaMemset(table, 0, (size_t)entries[0] * sizeof(float2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int t = 0; t < 2; ++t) {
        float best = 1e30f;
        for (int rep = 0; rep < 12; ++rep) {
            cudaEventRecord(e0);
            k_gather<<<(unsigned)((n_threads + 255) / 256), 256>>>(table, entries[t], out, n_threads);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        const double gathers = (double)n_threads * 8.0;
        printf("{\"table_entries\": %u, \"table_MB\": %.1f, \"gathers\": %.0f, \"us\": %.1f, \"Ggathers_per_s\": %.1f, \"GB_per_s_of_8B\": %.1f, \"GB_per_s_of_32B_sectors\": %.1f}\n",
               entries[t], entries[t] * 8.0 / 1e6, gathers, best * 1e3, gathers / (best * 1e-3) / 1e9, gathers * 8.0 / (best * 1e-3) / 1e9,
               gathers * 32.0 / (best * 1e-3) / 1e9);
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
