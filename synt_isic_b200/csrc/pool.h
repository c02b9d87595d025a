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
        void* p = nullptr;
        auto it = free_.find(bytes);
        if (it != free_.end() && !it->second.empty()) {
            p = it->second.back();
            it->second.pop_back();
        } else {
            SYNT_CUDA(cudaMalloc(&p, bytes));
            all_[p] = bytes;
            total_ += bytes;
        }
        if (journal_) journal_->push_back(p);
        return p;
    }
    void release(void* p) {
        if (!p) return;
        auto it = all_.find(p);
        SYNT_CHECK(it != all_.end(), "Pool::release of a foreign pointer");
        free_[it->second].push_back(p);
        if (journal_) {
            for (size_t i = journal_->size(); i-- > 0;)
                if ((*journal_)[i] == p) { journal_->erase(journal_->begin() + (long)i); break; }
        }
    }
    size_t total_bytes() const { return total_; }
    // Exception safety of a forward pass: while a Scope is alive every block handed out is journaled; if the scope is left
    // without commit() (an exception unwound through it) the blocks still out are returned to the pool, so a failed call does
    // not shrink the pool for the next one.
    class Scope {
    public:
        explicit Scope(Pool& p) : pool_(p), prev_(p.journal_) { p.journal_ = &mine_; }
        void commit() { done_ = true; }
        ~Scope() {
            pool_.journal_ = prev_;
            if (done_) { if (prev_) prev_->insert(prev_->end(), mine_.begin(), mine_.end()); return; }
            for (void* p : mine_) { auto it = pool_.all_.find(p); if (it != pool_.all_.end()) pool_.free_[it->second].push_back(p); }
        }
    private:
        Pool& pool_; std::vector<void*>* prev_; std::vector<void*> mine_; bool done_ = false;
    };
private:
    std::vector<void*>* journal_ = nullptr;
    std::map<void*, size_t> all_;
    std::map<size_t, std::vector<void*>> free_;
    size_t total_ = 0;
};

}  // namespace synt
