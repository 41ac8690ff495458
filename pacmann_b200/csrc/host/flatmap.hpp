// uint64 -> uint64 open-addressing hash map (linear probing, power-of-two capacity) for the per-call and per-search
// bookkeeping of the host mirror: responses by global index, local-cache slots, known vertices.  The node-based
// std::unordered_map it replaces cost ~100 ns and one allocation per insert, which at ~100 inserts per client and step was
// the largest host cost of a lock-step search step.  Keys may be any value except ~0 (kEmpty).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace pianopir {

class FlatMap {
public:
    static constexpr uint64_t kEmpty = ~0ull;
    // empty the map; capacity is kept (or grown) so that `expected` entries stay below half load
    void reset(size_t expected) {
        size_t cap = 16;
        while (cap < expected * 2 + 2) cap <<= 1;
        if (cap > keys.size()) {
            keys.assign(cap, kEmpty);
            vals.assign(cap, 0);
        } else {
            std::fill(keys.begin(), keys.end(), kEmpty);
        }
        used = 0;
    }
    size_t size() const { return used; }
    const uint64_t *find(uint64_t k) const {
        if (keys.empty()) return nullptr;
        const size_t mask = keys.size() - 1;
        for (size_t i = hash(k) & mask;; i = (i + 1) & mask) {
            if (keys[i] == k) return &vals[i];
            if (keys[i] == kEmpty) return nullptr;
        }
    }
    bool has(uint64_t k) const { return find(k) != nullptr; }
    // insert or overwrite
    void put(uint64_t k, uint64_t v) {
        if (keys.empty() || (used + 1) * 2 > keys.size()) grow();
        const size_t mask = keys.size() - 1;
        for (size_t i = hash(k) & mask;; i = (i + 1) & mask) {
            if (keys[i] == k) { vals[i] = v; return; }
            if (keys[i] == kEmpty) { keys[i] = k; vals[i] = v; used++; return; }
        }
    }

private:
    static size_t hash(uint64_t k) {
        k *= 0x9E3779B97F4A7C15ull;
        return (size_t)(k ^ (k >> 32));
    }
    void grow() {
        std::vector<uint64_t> ok, ov;
        ok.swap(keys);
        ov.swap(vals);
        const size_t cap = ok.empty() ? 16 : ok.size() * 2;
        keys.assign(cap, kEmpty);
        vals.assign(cap, 0);
        used = 0;
        for (size_t i = 0; i < ok.size(); i++)
            if (ok[i] != kEmpty) put(ok[i], ov[i]);
    }
    std::vector<uint64_t> keys, vals;
    size_t used = 0;
};

}  // namespace pianopir
