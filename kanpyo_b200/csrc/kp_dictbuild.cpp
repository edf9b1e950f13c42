// kp_dictbuild.cpp — host-side double-array construction for the product's dictionary builder.
//
// Produces the same base/check array as the reference's `da::build_with_ids`
// (kanpyo-dict/src/trie/da.rs:23-131,206-217), so that a dictionary built here is interchangeable with
// one built by the reference's `ipa-dict-builder`:
//   * array starts at 50*1024 nodes and doubles on demand            (da.rs:6-7,37-41)
//   * node 0's base is the moving search hint, advanced only when the scanned window is >= 95 % full
//                                                                     (da.rs:43-77)
//   * children are placed first-fit, the subtree is built depth-first in label order; a terminator
//     child (label 0) stores -id in its base                          (da.rs:79-131)
//   * trailing unused nodes (check == 0) are cut off                  (da.rs:29-35)
// Differences in mechanics only: an explicit work stack instead of recursion (IPADIC keys are up to
// 78 bytes deep, user dictionaries may be deeper) and key ranges instead of per-node branch vectors.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace {

struct DaNode {
    int32_t base, check;
};

struct Work {
    uint32_t node;    // index of the trie state in the array
    uint32_t depth;   // number of key bytes consumed
    uint64_t lo, hi;  // keys [lo, hi) share that prefix (keys are sorted)
};

class DaBuilder {
  public:
    DaBuilder(const uint8_t* blob, const uint64_t* off, const int64_t* ids) : blob_(blob), off_(off), ids_(ids) {
        nodes_.assign(50 * 1024, DaNode{0, 0});
        nodes_[0].base = 2;   // ROOT_ID + 1 (da.rs:25)
    }

    bool run(uint64_t n_keys) {
        std::vector<Work> stack;
        stack.push_back(Work{1, 0, 0, n_keys});
        uint8_t labels[256];
        uint64_t first[257];
        while (!stack.empty()) {
            Work w = stack.back();
            stack.pop_back();
            grow_to(w.node);
            // distinct next bytes of the keys in [lo, hi); a key that ends here contributes label 0
            int n_labels = 0;
            for (uint64_t k = w.lo; k < w.hi; k++) {
                uint64_t s = off_[k] + w.depth;
                uint8_t c = s < off_[k + 1] ? blob_[s] : 0;
                if (n_labels == 0 || labels[n_labels - 1] != c) {
                    labels[n_labels] = c;
                    first[n_labels] = k;
                    n_labels++;
                }
            }
            first[n_labels] = w.hi;
            const uint32_t base = place(labels, n_labels);
            nodes_[w.node].base = (int32_t)base;
            for (int i = 0; i < n_labels; i++) {
                uint32_t q = base + labels[i];
                if (nodes_[q].check != 0) return false;   // assert!, da.rs:111-116
                nodes_[q].check = (int32_t)w.node;
                if (labels[i] == 0) {
                    int64_t id = ids_[w.lo];              // ids[branches[0]], da.rs:120
                    if (id <= 0 || id > INT32_MAX) return false;
                    nodes_[q].base = -(int32_t)id;
                }
            }
            // depth-first, label order: push in reverse so the first child is expanded next
            for (int i = n_labels - 1; i >= 0; i--)
                if (labels[i] != 0) stack.push_back(Work{base + labels[i], w.depth + 1, first[i], first[i + 1]});
        }
        size_t len = nodes_.size();
        while (len > 1 && nodes_[len - 1].check == 0) len--;
        nodes_.resize(len);
        return true;
    }

    std::vector<DaNode> nodes_;

  private:
    void grow_to(uint64_t idx) {
        while (idx >= nodes_.size()) nodes_.resize(nodes_.size() * 2, DaNode{0, 0});
    }

    // first-fit base for a label set, starting at the hint kept in node 0 (da.rs:43-77)
    uint32_t place(const uint8_t* labels, int n) {
        const uint32_t hint = (uint32_t)nodes_[0].base;
        for (uint32_t cand = hint;; cand++) {
            grow_to(cand);
            bool fits = true;
            for (int i = 0; i < n; i++) {
                uint64_t q = (uint64_t)cand + labels[i];
                grow_to(q);
                if (nodes_[q].check != 0) {
                    fits = false;
                    break;
                }
            }
            if (!fits) continue;
            uint32_t used = 0;
            for (uint32_t x = hint; x <= cand; x++) used += nodes_[x].check != 0;
            if ((double)used / (double)(cand - hint + 1) >= 0.95) nodes_[0].base = (int32_t)cand + 1;
            return cand;
        }
    }

    const uint8_t* blob_;
    const uint64_t* off_;
    const int64_t* ids_;
};

}  // namespace

extern "C" {

// keys: n_keys sorted unique byte strings blob[off[k] .. off[k+1]); ids[k] = KeywordID stored at key k's leaf.
// On success *out is a malloc'ed int32 array of 2 * *out_len values ({base, check} per node).
int kp_da_build(const uint8_t* blob, const uint64_t* off, uint64_t n_keys, const int64_t* ids, int32_t** out,
                uint64_t* out_len) {
    if (!off || !out || !out_len || (n_keys && (!ids || (off[n_keys] && !blob)))) return -1;
    DaBuilder b(blob, off, ids);
    if (!b.run(n_keys)) return -3;
    size_t bytes = b.nodes_.size() * sizeof(DaNode);
    int32_t* p = (int32_t*)malloc(bytes ? bytes : 8);
    if (!p) return -5;
    memcpy(p, b.nodes_.data(), bytes);
    *out = p;
    *out_len = b.nodes_.size();
    return 0;
}

void kp_da_free(int32_t* p) { free(p); }

}  // extern "C"
