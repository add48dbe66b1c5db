// microbench.cu -- latencies that bound the event kernel's synchronisation round on B200:
//   * grid barrier over 128 CTAs x 128 threads: (a) one counter, everybody spins on it; (b) the same with
//     __nanosleep back-off; (c) one counter + one release flag per CTA written by the last arriver;
//   * dependent ld.global.cg chain (L2 hit latency);
//   * FP64: dependent DADD/DFMA chain, __ddiv_rn, __dsqrt_rn latency; DFMA throughput per SM (32 warps, ILP 8).
// Build + run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb tools/microbench.cu && /tmp/mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void barrier_kernel(unsigned *bar, unsigned *flags, int iters, long long *out) {
    unsigned epoch = 0;
    const int ncta = gridDim.x;
    long long t0 = 0;
    for (int it = 0; it < iters + 10; it++) {
        if (it == 10 && threadIdx.x == 0) t0 = clock64();
        __syncthreads();
        if (MODE < 2) {
            if (threadIdx.x == 0) {
                epoch += ncta;
                __threadfence();
                atomicAdd(bar, 1u);
                while ((int)(ld_acquire_u32(bar) - epoch) < 0) { if (MODE == 1) __nanosleep(40); }
            }
        } else if (MODE == 3) {
            // release folded into the arrival (no separate fence, no returned value), relaxed polling, one acquire fence
            if (threadIdx.x == 0) {
                epoch += ncta;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                unsigned v;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - epoch) < 0);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else if (MODE == 4) {
            // release arrival, acquire polling (no fences)
            if (threadIdx.x == 0) {
                epoch += ncta;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                while ((int)(ld_acquire_u32(bar) - epoch) < 0) {}
            }
        } else if (MODE == 5) {
            // fence + relaxed atomic arrival (as mode 0), relaxed polling, acquire fence at the end
            if (threadIdx.x == 0) {
                epoch += ncta;
                __threadfence();
                atomicAdd(bar, 1u);
                unsigned v;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - epoch) < 0);
                __threadfence();
            }
        } else {
            __shared__ int last;
            if (threadIdx.x == 0) {
                epoch += ncta;
                __threadfence();
                const unsigned old = atomicAdd(bar, 1u);
                last = (old + 1 == epoch);
            }
            __syncthreads();
            const unsigned gen = (unsigned)it + 1;
            if (last) {
                for (int c = threadIdx.x; c < ncta; c += blockDim.x)
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32 * c), "r"(gen) : "memory");
            }
            if (threadIdx.x == 0) {
                while (ld_acquire_u32(flags + 32 * blockIdx.x) != gen) {}
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

__global__ void l2_chain(const unsigned *idx, int iters, long long *out) {
    unsigned p = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) p = __ldcg(idx + p);
    long long t1 = clock64();
    out[0] = (t1 - t0) / iters;
    out[1] = p;
}

template <int OP>
__global__ void fp64_latency(double x, double y, int iters, long long *out, double *sink) {
    double a = x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (OP == 0) a = __dadd_rn(a, y);
        if (OP == 1) a = __fma_rn(a, y, y);
        if (OP == 2) a = __ddiv_rn(y, a);
        if (OP == 3) a = __dsqrt_rn(a) + y;
        if (OP == 4) a = __dmul_rn(a, y);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) * 100 / iters;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

__global__ void fp64_throughput(double x, double y, int iters, long long *out, double *sink) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = x + u;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = __fma_rn(a[u], y, y);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) s += a[u];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    unsigned *bar, *flags;
    long long *out, h[256];
    double *sink;
    cudaMalloc(&bar, 64);
    cudaMalloc(&flags, 256 * 128);
    cudaMalloc(&out, 256 * 8);
    cudaMalloc(&sink, 148 * 1024 * 8);
    const int iters = 2000;
    for (int ncta : {16, 64, 128, 148}) {
        for (int mode = 0; mode < 6; mode++) {
            cudaMemset(bar, 0, 64);
            cudaMemset(flags, 0, 256 * 128);
            void *args[] = {&bar, &flags, (void *)&iters, &out};
            const void *fn = mode == 0 ? (const void *)barrier_kernel<0> : mode == 1 ? (const void *)barrier_kernel<1>
                           : mode == 2 ? (const void *)barrier_kernel<2> : mode == 3 ? (const void *)barrier_kernel<3>
                           : mode == 4 ? (const void *)barrier_kernel<4> : (const void *)barrier_kernel<5>;
            cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(ncta), dim3(128), args, 0, 0);
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
            printf("grid barrier mode %d, %3d CTAs x 128 thr: %lld cycles  (%s)\n", mode, ncta, h[0], cudaGetErrorString(e));
        }
    }
    {
        const int n = 1 << 20;
        unsigned *hidx = new unsigned[n], *didx;
        for (int i = 0; i < n; i++) hidx[i] = (unsigned)(((long long)i * 7919 + 12345) % n);
        cudaMalloc(&didx, n * 4);
        cudaMemcpy(didx, hidx, n * 4, cudaMemcpyHostToDevice);
        l2_chain<<<1, 1>>>(didx, 4000, out);
        cudaDeviceSynchronize();
        l2_chain<<<1, 1>>>(didx, 4000, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("dependent ld.global.cg chain (4 MB footprint, L2): %lld cycles per load\n", h[0]);
    }
    const char *names[] = {"DADD", "DFMA", "DDIV (__ddiv_rn)", "DSQRT (__dsqrt_rn) + DADD", "DMUL"};
    for (int op = 0; op < 5; op++) {
        if (op == 0) fp64_latency<0><<<1, 32>>>(1.5, 1.000001, 4096, out, sink);
        if (op == 1) fp64_latency<1><<<1, 32>>>(1.5, 1.000001, 4096, out, sink);
        if (op == 2) fp64_latency<2><<<1, 32>>>(1.5, 1.000001, 4096, out, sink);
        if (op == 3) fp64_latency<3><<<1, 32>>>(1.5, 1.000001, 4096, out, sink);
        if (op == 4) fp64_latency<4><<<1, 32>>>(1.5, 1.000001, 4096, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("%-28s dependent latency: %.2f cycles\n", names[op], h[0] / 100.0);
    }
    for (int warps : {4, 8, 16, 32}) {
        const int it = 4096;
        fp64_throughput<<<148, warps * 32>>>(1.5, 1.000001, it, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
        const double fma_per_clk_sm = (double)warps * 32 * 8 * it / (double)mx;
        printf("DFMA throughput, %2d warps/SM, ILP 8: %.1f FMA/clk/SM\n", warps, fma_per_clk_sm);
    }
    {
        // wall-clock FP64 peak over the whole GPU
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int it = 1 << 16;
        fp64_throughput<<<148 * 2, 1024>>>(1.5, 1.000001, 256, out, sink);
        cudaEventRecord(e0);
        fp64_throughput<<<148 * 2, 1024>>>(1.5, 1.000001, it, out, sink);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 148 * 2 * 1024 * 8 * (double)it;
        printf("FP64 FMA peak (measured, whole GPU): %.2f TFLOP/s (%.3f ms)\n", flops / ms * 1e-9, ms);
    }
    return 0;
}
