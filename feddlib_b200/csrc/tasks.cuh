// Task programs of the block-task kernel k_task (star_kernels.cuh): one-time schedule, built with the gather maps.
//
// A TILE is a handful of consecutive row nodes of one bucket (at most kTaskMaxTets incident elements in total) that one
// warp assembles together.  Every nonzero node block (I, J) of a tile's rows is the sum of the local blocks of the
// elements that contain both nodes: a PAIR (m, jc) = (element slot m of the tile, canonical local column jc of that
// element).  A TASK is the work of one lane in one pass: up to kTaskQ pairs of ONE position, accumulated in registers.
// Positions with more pairs are split into a GROUP of tasks on adjacent lanes of the same pass (the partial sums are
// combined with shuffles, fixed order).  Groups are emitted in order of decreasing task size, so the lanes of a pass
// carry (nearly) equal work.  Every value is written exactly once; the summation order is fixed by the schedule.
//
// The scheduler is a plain function (host and device): the device kernel runs it with one thread per tile, the CPU
// unit test (tests/test_task_schedule.py via tests/cpp/task_sched_c.cpp) checks its invariants without a GPU.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define FB_HD __host__ __device__
#else
#define FB_HD
#endif

namespace fb {

constexpr int kTaskQ = 4;          // pairs per task
constexpr int kTaskMaxTets = 32;   // element slots per tile (5 bits)
constexpr int kTaskMaxNodes = 8;   // row nodes per tile (3 bits)
constexpr int kTaskMaxLen = 256;   // positions per node row (8 bits)
constexpr int kTaskMaxGroup = 8;   // tasks per group (ninc <= 32, Q = 4)

// task word: bits 0..35 pairs (9 bits each: m | jc << 5), 36..38 npairs, 39..46 position, 47..49 node slot,
// 50..52 rem (tasks of the same group on the following lanes), 53 head (this lane stores the block)
FB_HD inline uint64_t task_pack(const uint16_t *pairs, int npairs, int pos, int slot, int rem, int head)
{
    uint64_t w = 0;
    for (int s = 0; s < npairs; s++) w |= (uint64_t)(pairs[s] & 0x1ffu) << (9 * s);
    w |= (uint64_t)npairs << 36;
    w |= (uint64_t)pos << 39;
    w |= (uint64_t)slot << 47;
    w |= (uint64_t)rem << 50;
    w |= (uint64_t)head << 53;
    return w;
}

struct TaskTile {        // 16 bytes
    uint32_t q0;         // first row of the tile (index into the bucket-ordered row records)
    uint8_t n_nodes, n_tets, n_passes, flags;
    uint32_t task_off;   // first task of the tile, in units of 32 task words
    uint32_t pad;
};

// position, in its row, of canonical column node jc of an incidence record (records of kernels.cuh: 16-bit positions in
// the leading words)
FB_HD inline int task_rec_pos(const uint32_t *w, int jc) { return (int)((w[jc >> 1] >> (16 * (jc & 1))) & 0xffffu); }

// Schedules one tile.  Rows: n_nodes, lengths len[], incidence counts ninc[], records of node i at rec + k0[i]*rec_words
// (ninc[i] consecutive records), ncol canonical column nodes per element.  Returns the number of passes; with out != nullptr
// also writes 32 * passes task words.  Returns -1 if the tile violates the limits of the format.
FB_HD inline int schedule_tile(int n_nodes, const int *len, const int *ninc, const int64_t *k0, const uint32_t *rec, int rec_words,
                               int ncol, uint64_t *out)
{
    if (n_nodes < 1 || n_nodes > kTaskMaxNodes) return -1;
    int m0[kTaskMaxNodes + 1];
    m0[0] = 0;
    for (int i = 0; i < n_nodes; i++) {
        if (len[i] > kTaskMaxLen || ninc[i] < 0) return -1;
        m0[i + 1] = m0[i] + ninc[i];
    }
    if (m0[n_nodes] > kTaskMaxTets) return -1;
    int cursor = 0;
    for (int key = kTaskQ; key >= 0; key--) {
        for (int i = 0; i < n_nodes; i++) {
            const uint32_t *r = rec + k0[i] * rec_words;
            // counting sort of the node's pairs by position (pairs of a position keep the order (element, column))
            uint8_t cnt[kTaskMaxLen];
            for (int p = 0; p < len[i]; p++) cnt[p] = 0;
            for (int t = 0; t < ninc[i]; t++)
                for (int jc = 0; jc < ncol; jc++) {
                    const int p = task_rec_pos(r + (int64_t)t * rec_words, jc);
                    if (p >= len[i]) return -1;
                    cnt[p]++;
                }
            bool any = false;
            for (int p = 0; p < len[i] && !any; p++) {
                const int c = cnt[p], nt = c == 0 ? 1 : (c + kTaskQ - 1) / kTaskQ;
                any = (c + nt - 1) / nt == key;
            }
            if (!any) continue;
            uint16_t start[kTaskMaxLen + 1];
            uint16_t list[kTaskMaxTets * 10];
            if (out) {
                uint8_t fill[kTaskMaxLen];
                start[0] = 0;
                for (int p = 0; p < len[i]; p++) { start[p + 1] = (uint16_t)(start[p] + cnt[p]); fill[p] = 0; }
                for (int t = 0; t < ninc[i]; t++)
                    for (int jc = 0; jc < ncol; jc++) {
                        const int p = task_rec_pos(r + (int64_t)t * rec_words, jc);
                        list[start[p] + fill[p]++] = (uint16_t)((m0[i] + t) | (jc << 5));
                    }
            }
            for (int p = 0; p < len[i]; p++) {
                const int c = cnt[p], nt = c == 0 ? 1 : (c + kTaskQ - 1) / kTaskQ;
                if ((c + nt - 1) / nt != key) continue;   // size of the group's largest task
                if (nt > kTaskMaxGroup) return -1;
                if ((cursor & 31) + nt > 32) {            // a group never straddles two passes
                    while (cursor & 31) { if (out) out[cursor] = 0; cursor++; }
                }
                int done = 0;
                for (int j = 0; j < nt; j++) {
                    const int sz = c / nt + (j < c % nt ? 1 : 0);
                    if (out) out[cursor] = task_pack(list + start[p] + done, sz, p, i, nt - 1 - j, j == 0 ? 1 : 0);
                    done += sz;
                    cursor++;
                }
            }
        }
    }
    while (cursor & 31) { if (out) out[cursor] = 0; cursor++; }
    return cursor / 32;
}

} // namespace fb
