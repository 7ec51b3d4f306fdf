// How does the B200's block scheduler deal the CTAs of a sub-wave grid to the SMs?  Same launch
// shape as the raster (128 threads, ~26 KB static shared memory, 64 registers -> 8 CTAs per SM):
// every CTA records its SM and its start time, spins for a while (so that all CTAs are resident
// together, as in a single-wave raster launch), and the host prints which CTA indices each SM got.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_cta_dealing.bin tools/probe_cta_dealing.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 8) probe(int *smid, unsigned long long *t0, int spin)
{
    __shared__ float pad[6600];
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (threadIdx.x == 0) {
        smid[blockIdx.x] = (int)s;
        t0[blockIdx.x] = t;
    }
    float a = threadIdx.x;
    for (int i = 0; i < spin; ++i) a = a * 1.0001f + 0.5f;
    pad[threadIdx.x] = a;
    __syncthreads();
    if (pad[(threadIdx.x + 1) & 127] == 12345.f) smid[0] = -1;
}

int main(int argc, char **argv)
{
    const int grids[] = {296, 512, 768, 1024, 1184, 1536};
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int grid : grids) {
        int *d_smid;
        unsigned long long *d_t;
        cudaMalloc(&d_smid, grid * sizeof(int));
        cudaMalloc(&d_t, grid * sizeof(unsigned long long));
        for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128>>>(d_smid, d_t, 20000);
        cudaDeviceSynchronize();
        std::vector<int> smid(grid);
        std::vector<unsigned long long> t(grid);
        cudaMemcpy(smid.data(), d_smid, grid * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(t.data(), d_t, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        std::vector<std::vector<int>> per(sms);
        for (int i = 0; i < grid; ++i)
            if (smid[i] >= 0 && smid[i] < sms) per[smid[i]].push_back(i);
        printf("grid %d: first 40 CTAs -> SM:", grid);
        for (int i = 0; i < 40 && i < grid; ++i) printf(" %d", smid[i]);
        printf("\n   CTAs of SM 0..5:");
        for (int s = 0; s < 6; ++s) {
            printf(" [");
            for (int i : per[s]) printf("%d ", i);
            printf("]");
        }
        int mn = 1 << 30, mx = 0;
        for (int s = 0; s < sms; ++s) mn = std::min<int>(mn, per[s].size()), mx = std::max<int>(mx, per[s].size());
        // is the deal "CTA i -> f(i mod R)" for some round length R?  count CTAs whose SM equals the SM of CTA i - R
        for (int R : {sms, sms / 2, 2 * sms, 132, 144, 74}) {
            int same = 0, tot = 0;
            for (int i = R; i < grid; ++i) tot++, same += smid[i] == smid[i - R];
            if (tot) printf("\n   SM(i) == SM(i - %d) for %d of %d CTAs", R, same, tot);
        }
        printf("\n   CTAs per SM: min %d max %d; start-time spread %.1f us\n", mn, mx,
               (*std::max_element(t.begin(), t.end()) - *std::min_element(t.begin(), t.end())) / 1e3);
        cudaFree(d_smid);
        cudaFree(d_t);
    }
    return 0;
}
