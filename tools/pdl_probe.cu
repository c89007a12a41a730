// tools/pdl_probe.cu -- what does a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization wait for when
// the operation before it in the stream is NOT a kernel?  The library launches its stream kernels with that attribute and,
// when the host sees no data hazard with earlier KERNELS, lets them run without an initial griddepcontrol.wait.  This probe
// puts a host-to-device copy (and, in a second case, a copy that itself follows a kernel) in front of such a launch and
// checks whether the kernel sees the copied data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 tools/pdl_probe.cu -o tools/pdl_probe && tools/pdl_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_spin(float *p, size_t n, int iters) {   // a long-ish primary kernel
    asm volatile("griddepcontrol.launch_dependents;");
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = p[i];
    for (int k = 0; k < iters; ++k) v = v * 1.0000001f + 1e-9f;
    p[i] = v;
}
__global__ void k_read(const float *src, float *dst, size_t n, int wait_first) {
    asm volatile("griddepcontrol.launch_dependents;");
    if (wait_first) asm volatile("griddepcontrol.wait;" ::: "memory");
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
    if (!wait_first) asm volatile("griddepcontrol.wait;" ::: "memory");
}

static cudaError_t launch_read(cudaStream_t s, const float *src, float *dst, size_t n, int wait_first, bool attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (attr) { cfg.attrs = at; cfg.numAttrs = 1; }
    return cudaLaunchKernelEx(&cfg, k_read, src, dst, n, wait_first);
}

int main() {
    const size_t n = 64u << 20;   // 256 MiB: the copy takes ~5 ms over PCIe
    float *h, *d_src, *d_dst, *d_other, *h_out;
    CK(cudaMallocHost(&h, n * 4));
    CK(cudaMallocHost(&h_out, n * 4));
    CK(cudaMalloc(&d_src, n * 4));
    CK(cudaMalloc(&d_dst, n * 4));
    CK(cudaMalloc(&d_other, n * 4));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const char *names[] = {"copy -> kernel(attr, wait at END)", "copy -> kernel(attr, wait FIRST)", "copy -> kernel(plain launch)",
                           "kernel -> copy -> kernel(attr, wait at END)", "kernel -> copy -> kernel(attr, wait FIRST)"};
    for (int c = 0; c < 5; ++c) {
        long bad_total = 0;
        for (int rep = 0; rep < 6; ++rep) {
            const float tag = (float)(c * 100 + rep + 1);
            for (size_t i = 0; i < n; i += 4096) h[i] = tag;         // sampled positions carry the tag of this repetition
            CK(cudaMemsetAsync(d_src, 0, n * 4, s));
            CK(cudaMemsetAsync(d_dst, 0, n * 4, s));
            CK(cudaStreamSynchronize(s));
            if (c >= 3) k_spin<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_other, n, 200);
            CK(cudaMemcpyAsync(d_src, h, n * 4, cudaMemcpyHostToDevice, s));
            CK(launch_read(s, d_src, d_dst, n, (c == 1 || c == 4) ? 1 : 0, c != 2));
            CK(cudaMemcpyAsync(h_out, d_dst, n * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            long bad = 0;
            for (size_t i = 0; i < n; i += 4096) bad += h_out[i] != tag;
            bad_total += bad;
        }
        printf("{\"case\": \"%s\", \"stale_samples\": %ld, \"of\": %zu}\n", names[c], bad_total, 6 * (n / 4096));
    }
    return 0;
}
