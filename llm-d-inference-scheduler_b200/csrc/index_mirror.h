// index_mirror.h -- host-side write side of the prefix index: the reference's per-server LRU bookkeeping
// (approximateprefix/indexer.go:52-83 Add, :105-115 eviction callback, :167-182 RemovePod) kept in host memory;
// the device table (index_kernels.cu) is rebuilt from its (hash, endpoint) pairs on commit.
// LRU semantics are those of hashicorp/golang-lru v2.0.7 as used by the reference: Add of an existing key only
// refreshes recency; overflow evicts the oldest key and fires the eviction callback; Get never touches recency.
#pragma once
#include <cstdint>
#include <list>
#include <unordered_map>
#include <vector>

namespace epp {

class IndexMirror {
  public:
    explicit IndexMirror(int default_lru_size) : default_lru_(default_lru_size) {}

    void add(uint32_t server, const uint64_t *hashes, int n, int num_gpu_blocks) {
        auto it = lrus_.find(server);
        if (it == lrus_.end()) {                                     // indexer.go:57-69
            int size = num_gpu_blocks > 0 ? num_gpu_blocks : default_lru_;
            if (size <= 0) size = 1;
            it = lrus_.emplace(server, Lru{size, {}, {}}).first;
        }
        Lru &l = it->second;
        for (int i = 0; i < n; i++) {                                // indexer.go:71-74
            uint64_t h = hashes[i];
            auto p = l.pos.find(h);
            if (p != l.pos.end()) {
                l.order.splice(l.order.begin(), l.order, p->second);   // refresh recency
                continue;
            }
            l.order.push_front(h);
            l.pos[h] = l.order.begin();
            if ((int)l.order.size() > l.size) {
                uint64_t old = l.order.back();
                l.order.pop_back();
                l.pos.erase(old);
                evict(old, server);
            }
        }
        for (int i = 0; i < n; i++) {                                // indexer.go:76-83
            auto &set = hash_to_pods_[hashes[i]];
            bool found = false;
            for (uint32_t s : set) if (s == server) { found = true; break; }
            if (!found) { set.push_back(server); n_pairs_++; }
        }
        dirty_ = true;
    }

    void remove_pod(uint32_t server) {                               // indexer.go:167-182
        auto it = lrus_.find(server);
        if (it == lrus_.end()) return;
        for (auto rit = it->second.order.rbegin(); rit != it->second.order.rend(); ++rit) evict(*rit, server);
        lrus_.erase(it);
        dirty_ = true;
    }

    void clear() {
        hash_to_pods_.clear();
        lrus_.clear();
        n_pairs_ = 0;
        dirty_ = true;
    }

    void export_pairs(std::vector<uint64_t> &hashes, std::vector<uint32_t> &eps) const {
        hashes.clear();
        eps.clear();
        hashes.reserve(n_pairs_);
        eps.reserve(n_pairs_);
        for (const auto &kv : hash_to_pods_)
            for (uint32_t s : kv.second) { hashes.push_back(kv.first); eps.push_back(s); }
    }

    size_t n_pairs() const { return n_pairs_; }
    size_t n_hashes() const { return hash_to_pods_.size(); }
    bool dirty() const { return dirty_; }
    void mark_clean() { dirty_ = false; }
    bool empty() const { return hash_to_pods_.empty() && lrus_.empty(); }

  private:
    struct Lru {
        int size;
        std::list<uint64_t> order;                                   // front = most recent
        std::unordered_map<uint64_t, std::list<uint64_t>::iterator> pos;
    };
    void evict(uint64_t h, uint32_t server) {                        // makeEvictionFn, indexer.go:105-115
        auto it = hash_to_pods_.find(h);
        if (it == hash_to_pods_.end()) return;
        auto &set = it->second;
        for (size_t i = 0; i < set.size(); i++) {
            if (set[i] == server) {
                set[i] = set.back();
                set.pop_back();
                n_pairs_--;
                break;
            }
        }
        if (set.empty()) hash_to_pods_.erase(it);
    }
    int default_lru_;
    std::unordered_map<uint64_t, std::vector<uint32_t>> hash_to_pods_;
    std::unordered_map<uint32_t, Lru> lrus_;
    size_t n_pairs_ = 0;
    bool dirty_ = false;
};

}  // namespace epp
