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
