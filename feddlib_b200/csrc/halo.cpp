// Multi-rank host plan (include/feddb200_halo.h): ownership, Tpetra column map, ghost-row exchange plan, host-side
// globalAssemble and the unique -> repeated vector import.  Host code only; communication through the caller's callbacks.
#include "../../include/feddb200_halo.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <utility>
#include <vector>

#include "../../include/feddb200.h"

namespace fb {
void set_error(const std::string &msg);
}

#define HL_LOGIC(cond, msg)                                                                        \
    do {                                                                                           \
        if (cond) {                                                                                \
            fb::set_error(msg);                                                                    \
            return FEDDB200_ELOGIC;                                                                \
        }                                                                                          \
    } while (0)

struct feddb200_halo {
    int rank = 0, size = 1;
    int64_t nn = 0, n_owned = 0, n_ghost = 0, n_rows = 0, n_colmap = 0, n_cols = 0;
    std::vector<int64_t> gid_rep, owner;
    std::vector<int32_t> row_lid, col_lid, extra_row, extra_col;
    // final node pattern: the arrays of the pattern callback's last call (valid as long as the caller keeps them, see
    // feddb200_node_pattern_fn); not copied -- they are the largest arrays of the plan
    const int64_t *rowptr = nullptr;
    const int32_t *colind = nullptr;
    int64_t nnz_nodes = 0;
    std::vector<int64_t> unique_gids, ghost_row_gids, ghost_row_owner, colmap_gids, rep_of_row;
    std::vector<int64_t> send_counts_nodes, recv_counts_nodes;
    std::vector<int64_t> recv_row, recv_pos, recv_len_sender, recv_q;
    // import plan: owned rows each peer needs (in the peer's ghost order), my ghost rows per owner are contiguous
    std::vector<int64_t> imp_send_rows, imp_send_counts, imp_recv_counts;
};

namespace {

struct GidIndex {   // sorted (gid, index) pairs with binary search
    std::vector<std::pair<int64_t, int64_t>> v;
    void build(const int64_t *g, int64_t n)
    {
        v.resize((size_t)n);
        for (int64_t i = 0; i < n; i++) v[(size_t)i] = {g[i], i};
        std::stable_sort(v.begin(), v.end(), [](const std::pair<int64_t, int64_t> &a, const std::pair<int64_t, int64_t> &b) { return a.first < b.first; });
    }
    int64_t find(int64_t g) const
    {
        auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(g, (int64_t)INT64_MIN),
                                   [](const std::pair<int64_t, int64_t> &a, const std::pair<int64_t, int64_t> &b) { return a.first < b.first; });
        return (it != v.end() && it->first == g) ? it->second : -1;
    }
};

int exchange(const feddb200_comm *c, const std::vector<std::vector<int64_t>> &send, std::vector<int64_t> &flat, std::vector<int64_t> &counts)
{
    std::vector<int64_t> sc((size_t)c->size), sb;
    for (int d = 0; d < c->size; d++) { sc[(size_t)d] = (int64_t)send[(size_t)d].size(); sb.insert(sb.end(), send[(size_t)d].begin(), send[(size_t)d].end()); }
    counts.assign((size_t)c->size, 0);
    if (c->size == 1) { flat = sb; counts[0] = (int64_t)sb.size(); return FEDDB200_OK; }
    const int64_t *recv = nullptr;
    const int rc = c->alltoallv64(c->user, sb.data(), sc.data(), &recv, counts.data());
    HL_LOGIC(rc != 0, "halo plan: the communicator's alltoallv64 failed");
    const int64_t total = std::accumulate(counts.begin(), counts.end(), (int64_t)0);
    flat.assign(recv, recv + total);
    return FEDDB200_OK;
}

int factor_of(int rd, int cd, int mode) { return mode == FEDDB200_BLOCK_SCALAR ? 1 : (mode == FEDDB200_BLOCK_DIAG ? rd : rd * cd); }

} // namespace

extern "C" int feddb200_halo_create(feddb200_halo **out, const feddb200_comm *comm, int64_t nn, const int64_t *gid_rep, const int32_t *owner,
                                    feddb200_node_pattern_fn pattern_fn, void *pattern_user)
{
    HL_LOGIC(!out || !comm || !pattern_fn || nn < 0 || (nn > 0 && (!gid_rep || !owner)), "feddb200_halo_create: bad arguments");
    HL_LOGIC(comm->size > 1 && !comm->alltoallv64, "feddb200_halo_create: communicator without alltoallv64");
    feddb200_halo *H = new feddb200_halo;
    H->rank = comm->rank; H->size = comm->size; H->nn = nn;
    const int rank = comm->rank, size = comm->size;
    H->gid_rep.assign(gid_rep, gid_rep + nn);
    H->owner.resize((size_t)nn);
    for (int64_t i = 0; i < nn; i++) H->owner[(size_t)i] = owner[i];
    auto fail = [&](int rc) { delete H; return rc; };

    // 1. rows: owned nodes in repeated order (= unique-map order, Map_def.hpp:201-206), ghost nodes by (owner, gid)
    std::vector<int64_t> owned_ids, ghost_ids;
    for (int64_t i = 0; i < nn; i++) (owner[i] == rank ? owned_ids : ghost_ids).push_back(i);
    std::sort(ghost_ids.begin(), ghost_ids.end(), [&](int64_t a, int64_t b) {
        return owner[a] != owner[b] ? owner[a] < owner[b] : gid_rep[a] < gid_rep[b]; });
    H->n_owned = (int64_t)owned_ids.size(); H->n_ghost = (int64_t)ghost_ids.size(); H->n_rows = H->n_owned + H->n_ghost;
    H->row_lid.assign((size_t)nn, 0);
    H->rep_of_row.assign((size_t)H->n_rows, 0);
    for (int64_t k = 0; k < H->n_owned; k++) { H->row_lid[(size_t)owned_ids[(size_t)k]] = (int32_t)k; H->rep_of_row[(size_t)k] = owned_ids[(size_t)k]; }
    for (int64_t k = 0; k < H->n_ghost; k++) {
        H->row_lid[(size_t)ghost_ids[(size_t)k]] = (int32_t)(H->n_owned + k);
        H->rep_of_row[(size_t)(H->n_owned + k)] = ghost_ids[(size_t)k];
    }
    for (int64_t i : owned_ids) H->unique_gids.push_back(gid_rep[i]);
    for (int64_t i : ghost_ids) { H->ghost_row_gids.push_back(gid_rep[i]); H->ghost_row_owner.push_back(owner[i]); }

    // 2. preliminary pattern (columns = repeated ids) -> structure of the ghost rows to their owners
    const int64_t *rp0 = nullptr;
    const int32_t *ci0 = nullptr;
    int rc = pattern_fn(pattern_user, H->row_lid.data(), H->n_rows, H->n_owned, nullptr, nn, nullptr, nullptr, 0, &rp0, &ci0);
    if (rc != 0) return fail(rc);
    std::vector<int64_t> rec;   // (row gid, col gid, col owner) triples received
    std::vector<char> in_owned_rows((size_t)nn, 0);
    {
        std::vector<std::vector<int64_t>> send((size_t)size);
        for (int64_t r = H->n_owned; r < H->n_rows; r++) {
            const int64_t rep = H->rep_of_row[(size_t)r];
            std::vector<int64_t> &s = send[(size_t)owner[rep]];
            for (int64_t k = rp0[r]; k < rp0[r + 1]; k++) {
                const int64_t c = ci0[k];
                s.push_back(gid_rep[rep]); s.push_back(gid_rep[c]); s.push_back(owner[c]);
            }
        }
        for (int64_t k = 0; k < rp0[H->n_owned]; k++) in_owned_rows[(size_t)ci0[k]] = 1;
        std::vector<int64_t> counts;
        rc = exchange(comm, send, rec, counts);
        if (rc != FEDDB200_OK) return fail(rc);
    }
    const int64_t n_rec = (int64_t)rec.size() / 3;

    // 3. column index space: the Tpetra column map (owned, then remotes by (owner, gid)), then ghost-only columns
    GidIndex rep_index;
    rep_index.build(gid_rep, nn);
    std::vector<std::pair<int64_t, int64_t>> remote;   // (gid, owner)
    for (int64_t i = 0; i < nn; i++)
        if (owner[i] != rank && in_owned_rows[(size_t)i]) remote.push_back({gid_rep[i], owner[i]});
    for (int64_t k = 0; k < n_rec; k++)
        if (rec[(size_t)(3 * k + 2)] != rank) remote.push_back({rec[(size_t)(3 * k + 1)], rec[(size_t)(3 * k + 2)]});
    std::stable_sort(remote.begin(), remote.end(), [](const std::pair<int64_t, int64_t> &a, const std::pair<int64_t, int64_t> &b) { return a.first < b.first; });
    remote.erase(std::unique(remote.begin(), remote.end(), [](const std::pair<int64_t, int64_t> &a, const std::pair<int64_t, int64_t> &b) { return a.first == b.first; }),
                 remote.end());
    std::sort(remote.begin(), remote.end(), [](const std::pair<int64_t, int64_t> &a, const std::pair<int64_t, int64_t> &b) {
        return a.second != b.second ? a.second < b.second : a.first < b.first; });
    H->colmap_gids = H->unique_gids;
    for (const auto &g : remote) H->colmap_gids.push_back(g.first);
    H->n_colmap = (int64_t)H->colmap_gids.size();
    std::vector<int64_t> col_lid((size_t)nn, -1);
    for (int64_t k = 0; k < H->n_owned; k++) col_lid[(size_t)owned_ids[(size_t)k]] = k;
    for (size_t k = 0; k < remote.size(); k++) {
        const int64_t rep = rep_index.find(remote[k].first);
        if (rep >= 0) col_lid[(size_t)rep] = H->n_owned + (int64_t)k;
    }
    int64_t extra_cols = 0;
    for (int64_t i = 0; i < nn; i++)
        if (col_lid[(size_t)i] < 0) col_lid[(size_t)i] = H->n_colmap + extra_cols++;
    H->n_cols = H->n_colmap + extra_cols;
    H->col_lid.resize((size_t)nn);
    for (int64_t i = 0; i < nn; i++) H->col_lid[(size_t)i] = (int32_t)col_lid[(size_t)i];

    // 4. entries of the owned rows that only other ranks contribute
    GidIndex colmap_index;
    colmap_index.build(H->colmap_gids.data(), H->n_colmap);
    H->extra_row.resize((size_t)n_rec); H->extra_col.resize((size_t)n_rec);
    for (int64_t k = 0; k < n_rec; k++) {
        const int64_t rep = rep_index.find(rec[(size_t)(3 * k)]);
        if (rep < 0 || H->row_lid[(size_t)rep] >= H->n_owned) { fb::set_error("halo plan: ghost row sent to a rank that does not own it"); return fail(FEDDB200_ELOGIC); }
        const int64_t c = colmap_index.find(rec[(size_t)(3 * k + 1)]);
        if (c < 0) { fb::set_error("halo plan: received column missing from the column map"); return fail(FEDDB200_ELOGIC); }
        H->extra_row[(size_t)k] = H->row_lid[(size_t)rep];
        H->extra_col[(size_t)k] = (int32_t)c;
    }
    const int64_t *rp = nullptr;
    const int32_t *ci = nullptr;
    rc = pattern_fn(pattern_user, H->row_lid.data(), H->n_rows, H->n_owned, H->col_lid.data(), H->n_cols, H->extra_row.data(), H->extra_col.data(),
                    n_rec, &rp, &ci);
    if (rc != 0) return fail(rc);
    H->rowptr = rp; H->colind = ci; H->nnz_nodes = rp[H->n_rows];

    // 5. ghost rows in final CSR order -> the owners resolve every entry to (row, position)
    {
        std::vector<int64_t> gid_of_col((size_t)H->n_cols, -1);
        for (int64_t i = 0; i < nn; i++) gid_of_col[(size_t)H->col_lid[(size_t)i]] = gid_rep[i];
        for (int64_t k = 0; k < H->n_colmap; k++) gid_of_col[(size_t)k] = H->colmap_gids[(size_t)k];
        std::vector<std::vector<int64_t>> send((size_t)size);
        H->send_counts_nodes.assign((size_t)size, 0);
        for (int64_t r = H->n_owned; r < H->n_rows; r++) {
            const int64_t rep = H->rep_of_row[(size_t)r], len = H->rowptr[r + 1] - H->rowptr[r];
            std::vector<int64_t> &s = send[(size_t)owner[rep]];
            for (int64_t q = 0; q < len; q++) {
                s.push_back(gid_rep[rep]); s.push_back(gid_of_col[(size_t)H->colind[H->rowptr[r] + q]]); s.push_back(len); s.push_back(q);
            }
            H->send_counts_nodes[(size_t)owner[rep]] += len;
        }
        std::vector<int64_t> rec4, counts;
        rc = exchange(comm, send, rec4, counts);
        if (rc != FEDDB200_OK) return fail(rc);
        H->recv_counts_nodes.resize((size_t)size);
        for (int s = 0; s < size; s++) H->recv_counts_nodes[(size_t)s] = counts[(size_t)s] / 4;
        const int64_t n4 = (int64_t)rec4.size() / 4;
        H->recv_row.resize((size_t)n4); H->recv_pos.resize((size_t)n4); H->recv_len_sender.resize((size_t)n4); H->recv_q.resize((size_t)n4);
        for (int64_t k = 0; k < n4; k++) {
            const int64_t rep = rep_index.find(rec4[(size_t)(4 * k)]);
            const int64_t c = colmap_index.find(rec4[(size_t)(4 * k + 1)]);
            if (rep < 0 || c < 0) { fb::set_error("halo plan: received entry with an unknown row or column"); return fail(FEDDB200_ELOGIC); }
            const int64_t I = H->row_lid[(size_t)rep];
            const int32_t *b = H->colind + H->rowptr[I], *e = H->colind + H->rowptr[I + 1];
            const int32_t *it = std::lower_bound(b, e, (int32_t)c);
            if (it == e || *it != (int32_t)c) { fb::set_error("halo plan: received entry missing from the owner's pattern"); return fail(FEDDB200_ELOGIC); }
            H->recv_row[(size_t)k] = I; H->recv_pos[(size_t)k] = it - b;
            H->recv_len_sender[(size_t)k] = rec4[(size_t)(4 * k + 2)]; H->recv_q[(size_t)k] = rec4[(size_t)(4 * k + 3)];
        }
    }

    // import plan (unique -> repeated): every rank tells the owners which of their nodes it holds as ghosts
    {
        std::vector<std::vector<int64_t>> send((size_t)size);
        H->imp_recv_counts.assign((size_t)size, 0);
        for (int64_t k = 0; k < H->n_ghost; k++) {
            send[(size_t)H->ghost_row_owner[(size_t)k]].push_back(H->ghost_row_gids[(size_t)k]);
            H->imp_recv_counts[(size_t)H->ghost_row_owner[(size_t)k]]++;
        }
        std::vector<int64_t> want;
        rc = exchange(comm, send, want, H->imp_send_counts);
        if (rc != FEDDB200_OK) return fail(rc);
        H->imp_send_rows.resize(want.size());
        for (size_t k = 0; k < want.size(); k++) {
            const int64_t rep = rep_index.find(want[k]);
            if (rep < 0 || H->row_lid[(size_t)rep] >= H->n_owned) { fb::set_error("halo plan: a rank asked for a node this rank does not own"); return fail(FEDDB200_ELOGIC); }
            H->imp_send_rows[k] = H->row_lid[(size_t)rep];
        }
    }
    *out = H;
    return FEDDB200_OK;
}

extern "C" void feddb200_halo_free(feddb200_halo *H) { delete H; }

extern "C" int feddb200_halo_sizes(const feddb200_halo *H, int64_t *n_owned, int64_t *n_ghost, int64_t *n_rows, int64_t *n_colmap,
                                   int64_t *n_cols, int64_t *n_extra, int64_t *nnz_owned_nodes, int64_t *nnz_nodes, int64_t *n_recv)
{
    HL_LOGIC(!H, "feddb200_halo_sizes: null plan");
    if (n_owned) *n_owned = H->n_owned;
    if (n_ghost) *n_ghost = H->n_ghost;
    if (n_rows) *n_rows = H->n_rows;
    if (n_colmap) *n_colmap = H->n_colmap;
    if (n_cols) *n_cols = H->n_cols;
    if (n_extra) *n_extra = (int64_t)H->extra_row.size();
    if (nnz_owned_nodes) *nnz_owned_nodes = H->rowptr[H->n_owned];
    if (nnz_nodes) *nnz_nodes = H->nnz_nodes;
    if (n_recv) *n_recv = (int64_t)H->recv_row.size();
    return FEDDB200_OK;
}

extern "C" const void *feddb200_halo_array(const feddb200_halo *H, int which, int64_t *count)
{
    if (!H) return nullptr;
#define HL_ARR(v) do { if (count) *count = (int64_t)(v).size(); return (v).data(); } while (0)
    switch (which) {
    case 0: HL_ARR(H->row_lid);
    case 1: HL_ARR(H->col_lid);
    case 2: HL_ARR(H->extra_row);
    case 3: HL_ARR(H->extra_col);
    case 4: HL_ARR(H->colmap_gids);
    case 5: HL_ARR(H->unique_gids);
    case 6: HL_ARR(H->ghost_row_gids);
    case 7: HL_ARR(H->ghost_row_owner);
    case 8: if (count) *count = H->n_rows + 1; return H->rowptr;
    case 9: if (count) *count = H->nnz_nodes; return H->colind;
    case 10: HL_ARR(H->send_counts_nodes);
    case 11: HL_ARR(H->recv_counts_nodes);
    case 12: HL_ARR(H->recv_row);
    case 13: HL_ARR(H->recv_pos);
    case 14: HL_ARR(H->recv_len_sender);
    case 15: HL_ARR(H->recv_q);
    case 16: HL_ARR(H->imp_send_rows);
    case 17: HL_ARR(H->imp_send_counts);
    case 18: HL_ARR(H->rep_of_row);
    }
#undef HL_ARR
    return nullptr;
}

extern "C" int feddb200_halo_split_sizes(const feddb200_halo *H, int rd, int cd, int mode, int64_t *send_counts, int64_t *recv_counts)
{
    HL_LOGIC(!H, "feddb200_halo_split_sizes: null plan");
    const int f = factor_of(rd, cd, mode);
    for (int d = 0; d < H->size; d++) {
        if (send_counts) send_counts[d] = f * H->send_counts_nodes[(size_t)d];
        if (recv_counts) recv_counts[d] = f * H->recv_counts_nodes[(size_t)d];
    }
    return FEDDB200_OK;
}

// A sender ships each ghost node row as one block of f * len values in its own CSR order (a, q, b); the receiver adds
// element (a, q, b) to  f * base_I + a * per * L_I + per * p + b  (per = values per column node in a dof row).
extern "C" int feddb200_halo_recv_slots(const feddb200_halo *H, int rd, int cd, int mode, int64_t *slots)
{
    HL_LOGIC(!H || (!slots && !H->recv_row.empty()), "feddb200_halo_recv_slots: null argument");
    const int f = factor_of(rd, cd, mode);
    const int per = mode == FEDDB200_BLOCK_FULL ? cd : 1;
    const int nrow_dofs = mode == FEDDB200_BLOCK_SCALAR ? 1 : rd;
    const int64_t n = (int64_t)H->recv_row.size();
    int64_t block_start = 0, next_start = 0;
    for (int64_t k = 0; k < n; k++) {
        const int64_t Ls = H->recv_len_sender[(size_t)k], q = H->recv_q[(size_t)k];
        if (q == 0) { block_start = next_start; next_start += f * Ls; }
        const int64_t I = H->recv_row[(size_t)k], p = H->recv_pos[(size_t)k];
        const int64_t base = H->rowptr[I], L = H->rowptr[I + 1] - base;
        for (int a = 0; a < nrow_dofs; a++)
            for (int b = 0; b < per; b++) slots[block_start + a * per * Ls + per * q + b] = f * base + a * per * L + per * p + b;
    }
    return FEDDB200_OK;
}

extern "C" int feddb200_halo_export_add(const feddb200_halo *H, const feddb200_comm *comm, int rd, int cd, int mode, int64_t nnz_owned_values,
                                        double *values)
{
    HL_LOGIC(!H || !comm || !values, "feddb200_halo_export_add: null argument");
    if (H->size == 1) return FEDDB200_OK;
    static_assert(sizeof(double) == sizeof(int64_t), "doubles travel as 64-bit words");
    const int f = factor_of(rd, cd, mode);
    std::vector<int64_t> sc((size_t)H->size), rcnt((size_t)H->size);
    int64_t nsend = 0;
    for (int d = 0; d < H->size; d++) { sc[(size_t)d] = f * H->send_counts_nodes[(size_t)d]; nsend += sc[(size_t)d]; }
    (void)nsend;
    const int64_t *recv = nullptr;
    const int rc = comm->alltoallv64(comm->user, reinterpret_cast<const int64_t *>(values + nnz_owned_values), sc.data(), &recv, rcnt.data());
    HL_LOGIC(rc != 0, "halo export: the communicator's alltoallv64 failed");
    const int64_t nrecv = f * (int64_t)H->recv_row.size();
    std::vector<int64_t> slots((size_t)nrecv);
    const int rs = feddb200_halo_recv_slots(H, rd, cd, mode, slots.data());
    if (rs != FEDDB200_OK) return rs;
    for (int64_t t = 0; t < nrecv; t++) {
        double v;
        std::memcpy(&v, recv + t, sizeof(double));
        values[slots[(size_t)t]] += v;
    }
    return FEDDB200_OK;
}

extern "C" int feddb200_halo_import_vector(const feddb200_halo *H, const feddb200_comm *comm, int dofs, const double *u_unique, double *u_rep)
{
    HL_LOGIC(!H || !comm || !u_rep || (H->n_owned > 0 && !u_unique) || dofs < 1, "feddb200_halo_import_vector: bad arguments");
    for (int64_t k = 0; k < H->n_owned; k++)
        for (int d = 0; d < dofs; d++) u_rep[dofs * H->rep_of_row[(size_t)k] + d] = u_unique[dofs * k + d];
    if (H->size == 1) return FEDDB200_OK;
    std::vector<int64_t> send(H->imp_send_rows.size() * (size_t)dofs), sc((size_t)H->size), rcnt((size_t)H->size);
    for (size_t k = 0; k < H->imp_send_rows.size(); k++)
        std::memcpy(&send[k * (size_t)dofs], u_unique + dofs * H->imp_send_rows[k], sizeof(double) * (size_t)dofs);
    for (int d = 0; d < H->size; d++) sc[(size_t)d] = dofs * H->imp_send_counts[(size_t)d];
    const int64_t *recv = nullptr;
    const int rc = comm->alltoallv64(comm->user, send.data(), sc.data(), &recv, rcnt.data());
    HL_LOGIC(rc != 0, "halo import: the communicator's alltoallv64 failed");
    // my ghost rows are grouped by owner (ascending) = the arrival order
    for (int64_t k = 0; k < H->n_ghost; k++)
        std::memcpy(u_rep + dofs * H->rep_of_row[(size_t)(H->n_owned + k)], recv + dofs * k, sizeof(double) * (size_t)dofs);
    return FEDDB200_OK;
}
