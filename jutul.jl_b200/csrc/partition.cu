// Cell partitioning for block-Jacobi ILU(0) and for the multi-GPU decomposition.
// partition(N, num_coarse, weights; partitioner = MetisPartitioner()) src/partitioning.jl:244-307:
// graph vertices = cells, edges = faces, integer edge weights ceil(w / (0.1*mean(w))) clamped to
// [1, typemax(Int32)] (metis_integer_weights / generate_metis_graph :64-90), METIS k-way (:29-51).
// libmetis is a third-party binary in the reference too (Metis.jl); here the static build shipped with the
// CUDA toolkit is linked (idx_t = int64, probed). LinearPartitioner: partition_linear :12-18.
#include <algorithm>

#include "jb_internal.cuh"

extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                                   int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval,
                                   int64_t* part);

extern "C" {

int32_t jb_partition_linear(int64_t m, int64_t n, int64_t* part) {
    if (!part || m < 1 || n < 1) return JB_ERR_ARG;
    for (i64 i = 1; i <= n; i++) part[i - 1] = (i64)std::ceil((double)i / ((double)n / (double)m));
    return JB_OK;
}

int32_t jb_partition_metis(int64_t nc, int64_t nf, const int64_t* N, const double* weights, int64_t k, int64_t* part) {
    if (!N || !part || nc < 1 || nf < 0 || k < 1 || k > nc) return JB_ERR_ARG;
    if (k == 1) { for (i64 i = 0; i < nc; i++) part[i] = 1; return JB_OK; }
    if (k == nc) { for (i64 i = 0; i < nc; i++) part[i] = i + 1; return JB_OK; }
    std::vector<i64> w(nf, 1);
    if (weights) {
        double mean = 0.0;
        for (i64 f = 0; f < nf; f++) mean += weights[f];
        const double mv = (mean / (double)std::max<i64>(nf, 1)) * 0.1;
        for (i64 f = 0; f < nf; f++) w[f] = std::min<i64>(std::max<i64>((i64)std::ceil(weights[f] / mv), 1), 2147483647);
    }
    // symmetric adjacency; parallel faces between the same pair of cells merge with summed weights (sparse(I,J,V))
    std::vector<i64> xadj(nc + 1, 0);
    for (i64 f = 0; f < nf; f++) {
        i64 l = N[2 * f] - 1, r = N[2 * f + 1] - 1;
        if (l < 0 || l >= nc || r < 0 || r >= nc) return JB_ERR_ARG;
        if (l == r) continue;
        xadj[l + 1]++; xadj[r + 1]++;
    }
    for (i64 c = 0; c < nc; c++) xadj[c + 1] += xadj[c];
    std::vector<i64> adj(xadj[nc]), aw(xadj[nc]), cur(xadj.begin(), xadj.end() - 1);
    for (i64 f = 0; f < nf; f++) {
        i64 l = N[2 * f] - 1, r = N[2 * f + 1] - 1;
        if (l == r) continue;
        adj[cur[l]] = r; aw[cur[l]++] = w[f];
        adj[cur[r]] = l; aw[cur[r]++] = w[f];
    }
    std::vector<i64> xadj2(nc + 1, 0), adj2, aw2;
    adj2.reserve(adj.size()); aw2.reserve(adj.size());
    std::vector<std::pair<i64, i64>> row;
    for (i64 c = 0; c < nc; c++) {
        row.clear();
        for (i64 e = xadj[c]; e < xadj[c + 1]; e++) row.emplace_back(adj[e], aw[e]);
        std::sort(row.begin(), row.end());
        for (size_t e = 0; e < row.size(); e++) {
            if (!adj2.empty() && (i64)adj2.size() > xadj2[c] && adj2.back() == row[e].first) aw2.back() += row[e].second;
            else { adj2.push_back(row[e].first); aw2.push_back(row[e].second); }
        }
        xadj2[c + 1] = (i64)adj2.size();
    }
    i64 nvtxs = nc, ncon = 1, nparts = k, objval = 0;
    std::vector<i64> p(nc, 0);
    int rc = METIS_PartGraphKway(&nvtxs, &ncon, xadj2.data(), adj2.data(), nullptr, nullptr, aw2.data(), &nparts, nullptr, nullptr, nullptr,
                                 &objval, p.data());
    if (rc != 1) { jb_set_global_error("jb_partition_metis: METIS_PartGraphKway failed"); return JB_ERR_UNSUPPORTED; }
    std::vector<i64> cnt(k, 0);
    for (i64 c = 0; c < nc; c++) { part[c] = p[c] + 1; cnt[p[c]]++; }
    for (i64 b = 0; b < k; b++)
        if (cnt[b] == 0) { jb_set_global_error("Partitioning failed: a block is empty"); return JB_ERR_UNSUPPORTED; }
    return JB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Multicolour cell ordering for the GPU ILU(0) sweeps (setup-time, host). The reference lets the user renumber
// the local system before factorising (SymRCM option of the distributed path,
// ext/JutulPartitionedArraysExt/utils.jl:58-89); ILU(0) then eliminates in the new natural order. Here the
// order is chosen for the B200: cells are visited breadth-first from a pseudo-peripheral cell (Cuthill-McKee
// locality), greedily coloured in that order, and numbered colour by colour, breadth-first rank inside a colour.
// Rows of one colour are mutually independent, so the triangular sweeps have as many levels as colours (2 on
// bipartite hex-like grids) and every level is one contiguous, locality-ordered range of rows.
// perm[c_old] (1-based) = new label (1-based).
static int32_t order_multicolor_impl(int64_t nc, int64_t nf, const int64_t* N, const int64_t* last, int64_t* perm, int64_t* ncolors) {
    if (!N || !perm || nc < 1 || nf < 0) return JB_ERR_ARG;
    std::vector<i64> xadj(nc + 1, 0);
    for (i64 f = 0; f < nf; f++) {
        i64 l = N[2 * f] - 1, r = N[2 * f + 1] - 1;
        if (l < 0 || l >= nc || r < 0 || r >= nc) return JB_ERR_ARG;
        if (l == r) continue;
        xadj[l + 1]++; xadj[r + 1]++;
    }
    for (i64 c = 0; c < nc; c++) xadj[c + 1] += xadj[c];
    std::vector<int32_t> adj(xadj[nc]);
    {
        std::vector<i64> cur(xadj.begin(), xadj.end() - 1);
        for (i64 f = 0; f < nf; f++) {
            i64 l = N[2 * f] - 1, r = N[2 * f + 1] - 1;
            if (l == r) continue;
            adj[cur[l]++] = (int32_t)r; adj[cur[r]++] = (int32_t)l;
        }
    }
    std::vector<int32_t> order; order.reserve(nc);
    std::vector<int32_t> dist(nc, -1);
    auto bfs = [&](int32_t root, std::vector<int32_t>& out) {   // appends the component of root in BFS order
        size_t head = out.size();
        out.push_back(root); dist[root] = 0;
        while (head < out.size()) {
            int32_t v = out[head++];
            for (i64 e = xadj[v]; e < xadj[v + 1]; e++) {
                int32_t u = adj[e];
                if (dist[u] < 0) { dist[u] = dist[v] + 1; out.push_back(u); }
            }
        }
    };
    for (i64 s = 0; s < nc; s++) {
        if (dist[s] >= 0) continue;
        // pseudo-peripheral start: farthest cell of a first sweep
        std::vector<int32_t> comp;
        bfs((int32_t)s, comp);
        int32_t far = comp.back();
        for (int32_t v : comp) dist[v] = -1;
        bfs(far, order);
    }
    std::vector<int32_t> color(nc, -1);
    int32_t ncol = 0;
    std::vector<char> used;
    for (int32_t v : order) {
        used.assign(ncol + 1, 0);
        for (i64 e = xadj[v]; e < xadj[v + 1]; e++) {
            int32_t cu = color[adj[e]];
            if (cu >= 0) used[cu] = 1;
        }
        int32_t c = 0;
        while (c < ncol && used[c]) c++;
        color[v] = c;
        if (c == ncol) ncol++;
    }
    // numbering: colour by colour; inside a colour the cells flagged `last` (sub-domain boundary cells of a distributed
    // run) come after the others, each group in breadth-first rank
    std::vector<i64> start(2 * (size_t)ncol + 1, 0);
    auto bucket = [&](int32_t v) { return 2 * (size_t)color[v] + ((last && last[v]) ? 1 : 0); };
    for (i64 v = 0; v < nc; v++) start[bucket((int32_t)v) + 1]++;
    for (size_t b = 0; b < 2 * (size_t)ncol; b++) start[b + 1] += start[b];
    for (int32_t v : order) perm[v] = ++start[bucket(v)];     // 1-based
    if (ncolors) *ncolors = ncol;
    return JB_OK;
}
extern "C" int32_t jb_order_multicolor(int64_t nc, int64_t nf, const int64_t* N, int64_t* perm, int64_t* ncolors) {
    return order_multicolor_impl(nc, nf, N, nullptr, perm, ncolors);
}
// same, with the cells flagged in `last` (nc flags) numbered at the end of their colour: the interior rows of a rank
// then form contiguous ranges, so the interior SpMV can run while the halo is in flight
extern "C" int32_t jb_order_multicolor_boundary_last(int64_t nc, int64_t nf, const int64_t* N, const int64_t* last, int64_t* perm,
                                                      int64_t* ncolors) {
    return order_multicolor_impl(nc, nf, N, last, perm, ncolors);
}

// process_partition(neighbors, partition; weights) (src/partitioning.jl:128-160): every block of the partition that is
// not connected is split into its connected components; the first component keeps the label, the others receive
// max_p + 1, max_p + 2, ... in the order the blocks and components are met (components ordered by their smallest cell).
// Faces with zero weight do not connect cells. Setup-time host logic.
extern "C" int32_t jb_process_partition(int64_t nc, int64_t nf, const int64_t* N, const int64_t* partition, const double* weights,
                                        int64_t* new_partition) {
    if (!N || !partition || !new_partition || nc < 1 || nf < 0) return JB_ERR_ARG;
    i64 max_p = 0;
    for (i64 c = 0; c < nc; c++) { if (partition[c] < 1) return JB_ERR_ARG; new_partition[c] = partition[c]; max_p = std::max(max_p, partition[c]); }
    // block-internal adjacency (CSR)
    std::vector<i64> xadj(nc + 1, 0);
    auto internal = [&](i64 f, i64& l, i64& r) {
        l = N[2 * f] - 1; r = N[2 * f + 1] - 1;
        if (l < 0 || l >= nc || r < 0 || r >= nc) return false;
        if (weights && weights[f] == 0.0) return false;
        return partition[l] == partition[r] && l != r;
    };
    for (i64 f = 0; f < nf; f++) { i64 l, r; if (internal(f, l, r)) { xadj[l + 1]++; xadj[r + 1]++; } }
    for (i64 c = 0; c < nc; c++) xadj[c + 1] += xadj[c];
    std::vector<i64> adj(xadj[nc]), cur(xadj.begin(), xadj.end() - 1);
    for (i64 f = 0; f < nf; f++) { i64 l, r; if (internal(f, l, r)) { adj[cur[l]++] = r; adj[cur[r]++] = l; } }
    // cells grouped by block, ascending cell index inside a block
    std::vector<i64> bptr(max_p + 2, 0), bcells(nc);
    for (i64 c = 0; c < nc; c++) bptr[partition[c] + 1]++;
    for (i64 b = 1; b <= max_p + 1; b++) bptr[b] += bptr[b - 1];
    { std::vector<i64> bc(bptr.begin(), bptr.end() - 1); for (i64 c = 0; c < nc; c++) bcells[bc[partition[c]]++] = c; }
    std::vector<char> seen(nc, 0);
    std::vector<i64> stack;
    i64 next_label = max_p;
    for (i64 b = 1; b <= max_p; b++) {
        bool first = true;
        for (i64 k = bptr[b]; k < bptr[b + 1]; k++) {
            const i64 c = bcells[k];
            if (seen[c]) continue;
            i64 label = b;
            if (!first) label = ++next_label;
            first = false;
            stack.assign(1, c); seen[c] = 1;
            while (!stack.empty()) {
                const i64 u = stack.back(); stack.pop_back();
                new_partition[u] = label;
                for (i64 e = xadj[u]; e < xadj[u + 1]; e++) if (!seen[adj[e]]) { seen[adj[e]] = 1; stack.push_back(adj[e]); }
            }
        }
    }
    return JB_OK;
}
