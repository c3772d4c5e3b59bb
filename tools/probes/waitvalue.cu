// probe: cuStreamWaitValue32 through cudaGetDriverEntryPoint, on cudaMalloc and on pool memory
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <time.h>
typedef int (*WV)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
__global__ void setk(uint32_t *p, uint32_t v) { *p = v; __threadfence(); }
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    cudaFree(0);
    void *p1 = nullptr, *p2 = nullptr; cudaDriverEntryPointQueryResult r1, r2;
    cudaError_t e1 = cudaGetDriverEntryPoint("cuStreamWaitValue32", &p1, cudaEnableDefault, &r1);
    cudaError_t e2 = cudaGetDriverEntryPointByVersion("cuStreamWaitValue32", &p2, 12000, cudaEnableDefault, &r2);
    printf("default: err %d res %d ptr %p ; v12000: err %d res %d ptr %p\n", e1, r1, p1, e2, r2, p2);
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    uint32_t *m1, *m2; cudaMalloc(&m1, 256); cudaMallocAsync(&m2, 256, a); cudaMemsetAsync(m1, 0, 256, a); cudaMemsetAsync(m2, 0, 256, a); cudaStreamSynchronize(a);
    for (void *p : {p1, p2}) for (uint32_t *m : {m1, m2}) {
        if (!p) continue;
        int rc = ((WV)p)(b, (unsigned long long)(uintptr_t)m, 5, 0);
        printf("fn %p mem %p -> rc %d\n", p, (void *)m, rc);
        if (rc == 0) { setk<<<1, 1, 0, a>>>(m, 7); cudaStreamSynchronize(a); int spins = 0; cudaError_t e; while ((e = cudaStreamQuery(b)) == cudaErrorNotReady && spins < 2000) { spins++; struct timespec ts = {0, 1000000}; nanosleep(&ts, 0); } printf("  query b: %d after %d ms\n", e, spins); if (e == cudaErrorNotReady) { printf("  never released\n"); return 1; } cudaMemsetAsync(m, 0, 4, a); cudaStreamSynchronize(a); }
    }
    int v = 0; cudaDeviceGetAttribute(&v, (cudaDeviceAttr)92, 0); printf("attr92 %d\n", v);
    return 0;
}
