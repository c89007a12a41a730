// smb_alloc.h -- pooled storage for SMArray<T>::data.
//
// Replaces the reference's `new T[n]` / `delete[]` per array and per operator
// result (include/SMArray.h:41,62,219,342-346; include/UserFunctions.h:12,21,37,44).
// Every operator allocates a fresh result block (SMArray.h:219); a cudaMalloc /
// cudaFree pair per call (~100 us and a device-wide sync) would swamp a 2 us
// kernel, so freed blocks are kept in size-class free lists per (device, kind)
// and handed back without touching the driver.
//
// Views alias interior addresses of a block without refcounts
// (SMArray.h:128-135,427-436), so lookups accept any address inside a block.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <map>
#include <mutex>
#include <vector>

namespace smb {

struct Block {
    void *base;
    size_t bytes;   // rounded (bucket) size
    int device;     // owning device (-1 for pinned host)
    int kind;       // SMB_MEM_*
    // Managed blocks only: where the launchers last PUT the pages -- 0: unknown / maybe in host memory
    // (fresh from the driver, or the owner said the host wrote them); otherwise a signature of the
    // placement (one device, or the byte ranges a device set shares the block by).  A launcher that
    // needs placement P prefetches only when the recorded one differs, then records P; blocks last
    // written by a kernel / fill stay where they are.  A stale record is only a performance matter --
    // the GPU then demand-pages what moved.
    uint64_t placement = 0;
    // Managed blocks only: the block carries cudaMemAdviseSetReadMostly (it was last used as an operand several
    // devices read: the driver keeps a read-only duplicate on each, and invalidates them on any write -- the
    // host's raw writes through SMArray::data included).  Cleared, with the advice, when it becomes a result.
    bool read_mostly = false;
};

class Pool {
public:
    static Pool &instance() {
        static Pool p;
        return p;
    }
    // Returns nullptr on failure; *err carries the CUDA error.
    void *alloc(size_t bytes, int kind, int device, cudaError_t *err);
    bool free(void *ptr);
    bool owns(const void *ptr, Block *out = nullptr);
    // For a managed address: returns the containing block and whether its recorded placement
    // already was `want` (and records `want`).  false when the address is not in a live pool block.
    // rm_action: -1 leave the read-mostly mark, 0 clear it, 1 set it; *was_rm = the mark before.
    bool take_placement(const void *ptr, uint64_t want, Block *out, bool *matched, int rm_action = -1, bool *was_rm = nullptr);
    void clear_placement(const void *ptr);
    // Give cached blocks back to the driver until at most `keep_bytes` stay cached (largest first).
    void trim_to(uint64_t keep_bytes);
    void trim();
    void stats(uint64_t s[4]);

private:
    static size_t bucket(size_t bytes) {
        if (bytes == 0) bytes = 1;
        if (bytes <= (1u << 20)) return (bytes + 511) & ~size_t(511);          // 512 B steps below 1 MiB
        return (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);     // 2 MiB steps above
    }
    struct Key {
        int device, kind;
        size_t bytes;
        bool operator<(const Key &o) const {
            if (device != o.device) return device < o.device;
            if (kind != o.kind) return kind < o.kind;
            return bytes < o.bytes;
        }
    };
    cudaError_t raw_alloc(void **p, size_t bytes, int kind);
    void raw_free(const Block &b);

    std::mutex mu_;
    std::map<Key, std::vector<void *>> free_;
    std::map<uintptr_t, Block> live_;    // handed out, by base address
    std::map<uintptr_t, Block> cached_;  // sitting in free_
    uint64_t in_use_ = 0, cached_bytes_ = 0, driver_calls_ = 0, hits_ = 0;
};

} // namespace smb
