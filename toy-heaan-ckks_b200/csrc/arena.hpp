// arena.hpp -- address-ordered best-fit arena over a few large device segments (host-side bookkeeping only).
//
// Every stream-ordered allocation of 1 MiB and more made by the library is carved out of segments taken from the
// context's private cudaMemPool once.  The reference allocates a fresh Vec<[u64; N]> per polynomial and lets the
// system allocator cope; on the device, the driver's pool was measured to stall for 0.05 - 1 s whenever a loop of
// calls at changing levels (horner_chain: a different polynomial size at every level) makes it split, merge and
// re-grow.  All users of a context enqueue on one stream, so a range freed here is reusable at once in stream order.
//
// No CUDA in this file: segments come from a callback, which lets tests/emul exercise the bookkeeping on the CPU.
#pragma once
#include <cstddef>
#include <functional>
#include <iterator>
#include <map>
#include <unordered_map>
#include <vector>

class Arena {
public:
    static constexpr size_t ALIGN = (size_t)2 << 20;       // ranges are multiples of 2 MiB
    static constexpr size_t GROW_MIN = (size_t)64 << 20;   // first segment: 64 MiB (small contexts stay small) ...
    static constexpr size_t GROW_MAX = (size_t)2 << 30;    // ... then as large as everything held so far, up to 2 GiB
    struct Seg {
        char *base;
        size_t bytes;
    };
    // seg_alloc(bytes) returns a new segment or nullptr; seg_free(base) gives one back.
    typedef std::function<void *(size_t)> SegAlloc;
    typedef std::function<void(void *)> SegFree;

    // A range of at least `bytes`; *from_arena tells whether it was served without a new segment.  nullptr: out of memory.
    void *take(size_t bytes, const SegAlloc &seg_alloc, const SegFree &seg_free, bool *from_arena = nullptr) {
        const size_t need = (bytes + ALIGN - 1) / ALIGN * ALIGN;
        if (from_arena) *from_arena = true;
        if (void *p = carve(need)) return p;
        if (from_arena) *from_arena = false;
        size_t grow = total_ < GROW_MIN ? GROW_MIN : (total_ > GROW_MAX ? GROW_MAX : total_);
        size_t seg = need > grow ? need : grow;
        seg = (seg + GROW_MIN - 1) / GROW_MIN * GROW_MIN;  // whole multiples of 64 MiB
        void *base = seg_alloc(seg);
        if (!base && seg > need) {  // no room for a rounded segment: exactly what is needed
            seg = need;
            base = seg_alloc(seg);
        }
        if (!base) {  // give the entirely free segments back and try once more
            release_free_segments(seg_free);
            base = seg_alloc(seg);
            if (!base) return nullptr;
        }
        segs_.push_back({(char *)base, seg});
        total_ += seg;
        free_.emplace((char *)base, seg);
        return carve(need);
    }
    // true if p was handed out by take() (and is now free again); false: not ours.
    bool give(void *p) {
        auto it = live_.find(p);
        if (it == live_.end()) return false;
        char *base = (char *)p;
        size_t sz = it->second;
        live_.erase(it);
        char *seg_lo = nullptr, *seg_hi = nullptr;
        for (const Seg &sg : segs_)
            if (base >= sg.base && base < sg.base + sg.bytes) {
                seg_lo = sg.base;
                seg_hi = sg.base + sg.bytes;
            }
        auto nx = free_.lower_bound(base);
        if (nx != free_.end() && nx->first == base + sz && nx->first < seg_hi) {  // merge with the next free range
            sz += nx->second;
            nx = free_.erase(nx);
        }
        if (nx != free_.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == base && pv->first >= seg_lo) {  // merge with the previous one
                base = pv->first;
                sz += pv->second;
                free_.erase(pv);
            }
        }
        free_.emplace(base, sz);
        return true;
    }
    // Segments without a live range go back through seg_free.
    void release_free_segments(const SegFree &seg_free) {
        for (size_t i = 0; i < segs_.size();) {
            auto it = free_.find(segs_[i].base);
            if (it != free_.end() && it->second == segs_[i].bytes) {
                seg_free(segs_[i].base);
                total_ -= segs_[i].bytes;
                free_.erase(it);
                segs_[i] = segs_.back();
                segs_.pop_back();
            } else {
                ++i;
            }
        }
    }
    size_t total_bytes() const { return total_; }
    size_t live_ranges() const { return live_.size(); }
    size_t free_ranges() const { return free_.size(); }
    size_t segments() const { return segs_.size(); }
    size_t free_bytes() const {
        size_t s = 0;
        for (const auto &kv : free_) s += kv.second;
        return s;
    }

private:
    void *carve(size_t need) {  // best fit
        auto best = free_.end();
        for (auto it = free_.begin(); it != free_.end(); ++it)
            if (it->second >= need && (best == free_.end() || it->second < best->second)) best = it;
        if (best == free_.end()) return nullptr;
        char *base = best->first;
        const size_t sz = best->second;
        free_.erase(best);
        if (sz > need) free_.emplace(base + need, sz - need);
        live_[base] = need;
        return base;
    }
    std::vector<Seg> segs_;
    std::map<char *, size_t> free_;             // free ranges by address, never spanning two segments
    std::unordered_map<void *, size_t> live_;   // ranges handed out
    size_t total_ = 0;
};
