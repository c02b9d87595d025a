// Stream-ordered caching pool for activation buffers.  Every buffer is used on ONE stream,
// so a block may be handed out again as soon as it is released (stream order protects it).
// After the first forward of a given batch size no cudaMalloc happens any more, which is
// what makes the step capturable into a CUDA graph.
#pragma once
#include "common.cuh"
#include <map>
#include <vector>

namespace synt {

class Pool {
public:
    ~Pool() { for (auto& kv : all_) cudaFree(kv.first); }
    void* alloc(size_t bytes) {
        bytes = (bytes + 1023) & ~size_t(1023);
        auto it = free_.find(bytes);
        if (it != free_.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            return p;
        }
        void* p = nullptr;
        SYNT_CUDA(cudaMalloc(&p, bytes));
        all_[p] = bytes;
        total_ += bytes;
        return p;
    }
    void release(void* p) {
        if (!p) return;
        auto it = all_.find(p);
        SYNT_CHECK(it != all_.end(), "Pool::release of a foreign pointer");
        free_[it->second].push_back(p);
    }
    size_t total_bytes() const { return total_; }
private:
    std::map<void*, size_t> all_;
    std::map<size_t, std::vector<void*>> free_;
    size_t total_ = 0;
};

}  // namespace synt
