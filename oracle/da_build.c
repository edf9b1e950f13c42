/* TEST INFRASTRUCTURE (oracle) — not product code.
 *
 * Plain-C restatement of the reference's double-array construction,
 * kanpyo-dict/src/trie/da.rs:
 *   DoubleArray::new      :23-27   (zeroed nodes, nodes[0].base = ROOT_ID + 1)
 *   truncate              :29-35
 *   expand                :37-41   (length doubles)
 *   seek                  :43-77   (first-fit from the nodes[0].base hint; hint advances only when
 *                                   the scanned window is >= 95 % occupied)
 *   add                   :79-131  (depth-first, children in ascending byte order,
 *                                   leaf.base = -id on the TERMINATOR(0) edge)
 *   build / build_with_ids:191-217
 * The arrays it produces must be bit-identical to the reference's, because the dictionary file
 * stores them verbatim (da.rs:219-246) and the product's own builder is checked against them.
 *
 * Keys arrive as one byte blob + offsets; they must be sorted bytewise and unique (index.rs:16-38
 * guarantees that for the reference).  Because they are sorted, every `branches` vector the
 * reference passes down is a contiguous index range, so ranges [lo,hi) replace the Vec<KeywordID>.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define INIT_BUFFER_SIZE (50 * 1024) /* da.rs:6 */
#define EXPAND_RATIO 2               /* da.rs:7 */
#define TERMINATOR 0                 /* da.rs:8 */
#define ROOT_ID 1                    /* da.rs:9 */

typedef struct {
    int32_t base;
    int32_t check;
} ko_da_node;

typedef struct {
    ko_da_node *nodes;
    size_t len;
    const uint8_t *blob;
    const uint64_t *off;
    const int64_t *ids;
    int oom;
} builder;

static void expand(builder *b) { /* da.rs:37-41 */
    size_t new_len = b->len * EXPAND_RATIO;
    ko_da_node *n = (ko_da_node *)realloc(b->nodes, new_len * sizeof(ko_da_node));
    if (!n) {
        b->oom = 1;
        abort();
    }
    memset(n + b->len, 0, (new_len - b->len) * sizeof(ko_da_node));
    b->nodes = n;
    b->len = new_len;
}

static size_t seek(builder *b, const uint8_t *chars, int n_chars) { /* da.rs:43-77 */
    size_t left = (size_t)b->nodes[0].base;
    for (size_t i = left;; i++) {
        int found = 1;
        while (i >= b->len) expand(b);
        for (int k = 0; k < n_chars; k++) {
            int32_t q = (int32_t)i + (int32_t)chars[k];
            while (q >= (int32_t)b->len) expand(b);
            if (b->nodes[q].check != 0) {
                found = 0;
                break;
            }
        }
        if (found) {
            size_t used = 0;
            for (size_t x = left; x <= i; x++)
                if (b->nodes[x].check != 0) used++;
            double occupancy = (double)used / (double)(i - left + 1);
            if (occupancy >= 0.95) b->nodes[0].base = (int32_t)i + 1;
            return i;
        }
    }
}

static inline int key_byte(const builder *b, uint64_t k, size_t i) { /* str.get(i).unwrap_or(&TERMINATOR), da.rs:96 */
    uint64_t s = b->off[k], e = b->off[k + 1];
    return (s + i < e) ? b->blob[s + i] : TERMINATOR;
}

static void add(builder *b, size_t p, size_t depth, uint64_t lo, uint64_t hi) { /* da.rs:79-131 */
    while (p >= b->len) expand(b);
    uint8_t chars[256];
    uint64_t child_lo[256], child_hi[256];
    int n_chars = 0;
    for (uint64_t k = lo; k < hi; k++) {
        int ch = key_byte(b, k, depth);
        if (n_chars == 0 || chars[n_chars - 1] != (uint8_t)ch) {
            chars[n_chars] = (uint8_t)ch;
            child_lo[n_chars] = k;
            child_hi[n_chars] = k;
            n_chars++;
        }
        child_hi[n_chars - 1] = k + 1;
    }
    size_t left = seek(b, chars, n_chars);
    b->nodes[p].base = (int32_t)left;
    for (int c = 0; c < n_chars; c++) {
        int32_t q = (int32_t)left + chars[c];
        if (b->nodes[q].check != 0) abort(); /* assert!, da.rs:111-116 */
        b->nodes[q].check = (int32_t)p;
        if (chars[c] == TERMINATOR) {
            int32_t idx = -(int32_t)b->ids[lo]; /* ids[branches[0]], da.rs:120 */
            if (idx >= 0) abort();
            b->nodes[q].base = idx;
        }
    }
    for (int c = 0; c < n_chars; c++) {
        if (chars[c] == TERMINATOR) continue; /* TERMINATOR is never inserted in the child map, da.rs:101 */
        int32_t q = b->nodes[p].base + chars[c];
        add(b, (size_t)q, depth + 1, child_lo[c], child_hi[c]);
    }
}

/* build_with_ids (da.rs:206-217). Returns a malloc'ed array of *out_len nodes (free with ko_da_free). */
int ko_da_build(const uint8_t *blob, const uint64_t *off, uint64_t n_keys, const int64_t *ids,
                ko_da_node **out, uint64_t *out_len) {
    builder b;
    b.len = INIT_BUFFER_SIZE;
    b.nodes = (ko_da_node *)calloc(b.len, sizeof(ko_da_node));
    if (!b.nodes) return -1;
    b.nodes[0].base = ROOT_ID + 1; /* da.rs:25 */
    b.blob = blob;
    b.off = off;
    b.ids = ids;
    b.oom = 0;
    add(&b, ROOT_ID, 0, 0, n_keys);
    /* truncate, da.rs:29-35 */
    size_t len = b.len;
    while (len > 1 && b.nodes[len - 1].check == 0) len--;
    *out = b.nodes;
    *out_len = len;
    return 0;
}

void ko_da_free(ko_da_node *p) { free(p); }
