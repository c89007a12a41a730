// tools/overhead.cpp -- per-call host overhead of the C ABI (no Python in the loop).
//   g++ -O2 tools/overhead.cpp -Iinclude -Lsimplemath_b200 -lsmb200 -Wl,-rpath,'$ORIGIN/../simplemath_b200' -o tools/overhead
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <smb200.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const uint64_t n = 1024;
    for (int kind = 0; kind < 2; ++kind) {
        float *a = (float *)smb_alloc(n * 4, kind), *b = (float *)smb_alloc(n * 4, kind), *o = (float *)smb_alloc(n * 4, kind);
        if (!a) { printf("alloc failed: %s\n", smb_last_error()); return 1; }
        float one = 1.f;
        smb_fill(SMB_F32, a, &one, n, nullptr);
        smb_fill(SMB_F32, b, &one, n, nullptr);
        const int reps = 50000;
        for (int i = 0; i < 1000; ++i) smb_contiguous(SMB_OP_ADD, SMB_F32, a, b, o, n, nullptr);
        double t0 = now();
        for (int i = 0; i < reps; ++i) smb_contiguous(SMB_OP_ADD, SMB_F32, a, b, o, n, nullptr);
        double t1 = now();
        printf("kind %d: smb_contiguous n=1024 synchronous: %.2f us/call\n", kind, (t1 - t0) / reps * 1e6);
        float two = 2.f;
        t0 = now();
        for (int i = 0; i < reps; ++i) smb_array_scalar(SMB_OP_MUL, SMB_F32, a, &two, n, o, nullptr);
        t1 = now();
        printf("kind %d: smb_array_scalar n=1024 synchronous: %.2f us/call\n", kind, (t1 - t0) / reps * 1e6);
        t0 = now();
        for (int i = 0; i < reps; ++i) { void *p = smb_alloc(4 << 20, kind); smb_free(p); }
        t1 = now();
        printf("kind %d: smb_alloc+smb_free 4 MiB (pool hit): %.2f us/pair\n", kind, (t1 - t0) / reps * 1e6);
        const uint64_t m = 1000000;
        float *x = (float *)smb_alloc(m * 4, kind), *y = (float *)smb_alloc(m * 4, kind);
        smb_fill(SMB_F32, x, &one, m, nullptr);
        smb_fill(SMB_F32, y, &one, m, nullptr);
        t0 = now();
        for (int i = 0; i < 20000; ++i) { float *r = (float *)smb_alloc(m * 4, kind); smb_contiguous(SMB_OP_ADD, SMB_F32, x, y, r, m, nullptr); smb_free(r); }
        t1 = now();
        printf("kind %d: million_check as shipped (alloc + a+b + free, synchronous): %.2f us/iter\n", kind, (t1 - t0) / 20000 * 1e6);
    }
    return 0;
}
