// read_ceiling.cu — how fast can sm_100a pull bytes from HBM into shared memory with the access pattern of
// the stream kernels (cp.async.bulk chunks into an mbarrier ring, 2 CTAs per SM), with NO compute at all?
// The result is the practical ceiling the SpMV kernels are measured against besides the copy figure of
// MEASURED_PEAKS.json (a copy moves read+write; an SpMV is ~99 % reads).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o read_ceiling read_ceiling.cu && ./read_ceiling
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int kStages = 4, kChunk = 20480, kThreads = 288;

__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kThreads, 2) read_kernel(const unsigned char *src, size_t bytes_per_cta, int *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + kStages * kChunk);
    uint64_t *empty = full + kStages;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&full[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&empty[i])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char *base = src + (size_t)blockIdx.x * bytes_per_cta;
    const uint32_t nchunks = (uint32_t)(bytes_per_cta / kChunk);
    auto wait = [](uint64_t *bar, uint32_t parity) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                         su32(bar)),
                     "r"(parity)
                     : "memory");
    };
    if (threadIdx.x == 256) {           // producer
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        for (uint32_t q = 0; q < nchunks; ++q) {
            const uint32_t s = q % kStages;
            if (q >= kStages) wait(&empty[s], ((q / kStages) - 1) & 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full[s])), "r"(kChunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                             su32(sm + s * kChunk)),
                         "l"(base + (size_t)q * kChunk), "r"(kChunk), "r"(su32(&full[s])), "l"(pol)
                         : "memory");
        }
    } else if (threadIdx.x < 256) {     // consumers: wait, touch one word, release
        int acc = 0;
        for (uint32_t q = 0; q < nchunks; ++q) {
            const uint32_t s = q % kStages;
            wait(&full[s], (q / kStages) & 1);
            acc += reinterpret_cast<const int *>(sm + s * kChunk)[threadIdx.x];
            __syncwarp();
            if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&empty[s])) : "memory");
        }
        if (acc == 0x7fffffff) *sink = acc;
    }
}

int main() {
    const size_t total = (size_t)12 << 30;            // 12 GiB, far beyond the 126 MB L2
    unsigned char *buf;
    int *sink;
    cudaMalloc(&buf, total);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, total);
    const int smem = kStages * kChunk + 2 * kStages * 8;
    cudaFuncSetAttribute(read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int ctas : {296, 592, 2368, 9472}) {
        const size_t per = total / ctas / kChunk * kChunk;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0);
            read_kernel<<<ctas, kThreads, smem>>>(buf, per, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        printf("{\"microbench\": \"tma_bulk_read\", \"ctas\": %d, \"bytes\": %zu, \"ms\": %.4f, \"GBps\": %.1f, \"err\": \"%s\"}\n", ctas,
               per * ctas, best, per * ctas / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
