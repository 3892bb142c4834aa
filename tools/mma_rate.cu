// mma_rate.cu — measures the issue rate of tcgen05.mma (cta_group::1, M=128, N=256, SS operands) for kind::i8, kind::f16
// (bf16) and kind::f8f6f4 (e4m3) on this GPU: cycles per instruction and MAC/clk/SM.  Operand contents are irrelevant
// (zero-filled shared memory); every CTA issues `iters` back-to-back MMAs into one TMEM accumulator and waits for the
// final commit.  Used to pin the tensor-pipe roofline for the int8 path (MEASURED_PEAKS.json only has a bf16 figure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}

template <int KIND>  // 0 = i8, 1 = f16 (bf16 inputs, f32 accum), 2 = f8f6f4 (e4m3, f32 accum)
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int n_dim, long long* out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (128 + 256) * 128; i += blockDim.x) smem[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_ptr;
    long long cycles = 0;
    if (threadIdx.x == 0) {
        // idesc: c_format [4,6), a_format [7,10), b_format [10,13), n>>3 [17,23), m>>4 [24,29)
        uint32_t idesc;
        if (KIND == 0) idesc = (2u << 4) | (0u << 7) | (0u << 10);            // s32 <- u8 x u8
        else if (KIND == 1) idesc = (1u << 4) | (1u << 7) | (1u << 10);       // f32 <- bf16 x bf16
        else idesc = (1u << 4) | (0u << 7) | (0u << 10);                      // f32 <- e4m3 x e4m3
        idesc |= ((uint32_t)(n_dim >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t a_desc = make_desc(smem_u32(smem), 1024, 2);
        const uint64_t b_desc = make_desc(smem_u32(smem + 128 * 128), 1024, 2);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (KIND == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
            else if (KIND == 1)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 4000000000ll) break;
        }
        cycles = clock64() - t0;
        out_cycles[blockIdx.x] = cycles;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int KIND>
void run(const char* name, int k_elems, int n_dim, int iters, int sms) {
    long long* d;
    cudaMalloc(&d, sms * sizeof(long long));
    const size_t smem = (128 + 256) * 128;
    cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rate_kernel<KIND><<<sms, 128, smem>>>(iters, n_dim, d);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[256]; cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
        const double macs = 128.0 * n_dim * k_elems;
        if (rep == 1)
            printf("{\"kind\": \"%s\", \"m\": 128, \"n\": %d, \"k\": %d, \"ctas\": %d, \"iters\": %d, \"cycles_per_mma\": %.1f, \"mac_per_clk_per_sm\": %.0f, \"kernel_ms\": %.3f, "
                   "\"chip_tera_ops_per_s\": %.0f}\n", name, n_dim, k_elems, sms, iters, avg / iters, macs * iters / avg, ms, 2.0 * macs * iters * sms / (ms * 1e-3) / 1e12);
    }
    cudaFree(d);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    for (int n : {256, 208, 128}) {
        run<0>("i8", 32, n, 20000, sms);
        run<1>("bf16", 16, n, 20000, sms);
        run<2>("e4m3", 32, n, 20000, sms);
    }
    run<0>("i8", 32, 256, 20000, 1);
    return 0;
}
