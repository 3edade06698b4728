// Tree gauge on the device: gauge_walk(tn, edges) (src/abstractitensornetwork.jl:387-393), the loop of qr!(tn, edge)
// behind tree_gauge / tree_orthogonalize (:395-420) that apply(o, psi; ortho = true) calls before a gate
// (src/apply.jl:109-111, 130-132).
//
// One step "qr!(tn, src => dst)" replaces the tensor of `src` by an isometry Q over (site, other bonds) -> bond and
// multiplies the factor R into `dst`.  Only Q R = A matters to every caller (the gauge is not unique: the reference's
// Householder R is upper triangular, and the reference's own tests compare gauge-invariant quantities), so the step is
// done with the kernels the simple update already has instead of a Householder sweep over a tall matrix:
//
//   G  = A^H A over everything but the bond      one DMMA close per edge (itn_run_vertex_jobs, no messages absorbed)
//   R  = G^(1/2),  R^+ = G^(-1/2)                batched Hermitian Jacobi (itn_dev_map_eigvals; eigenvalues below
//                                                10 eps tr G are dropped: a rank-deficient bond gets a partial isometry)
//   Q  = A x_bond R^+,   B' = R x_bond B         DMMA mode products (itn_run_modeprods)
//
// R^+ from the Gram matrix carries kappa(A)^2 eps; a second pass over Q (kappa = 1 + O(kappa^2 eps)) restores
// Q^H Q = 1 to eps as in CholeskyQR2, R = R_2 R_1 being applied to B factor by factor.
// Edges of the sequence whose sources were not touched by an earlier edge of the same level are batched: one launch
// sequence per level of the tree, not per edge.
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>

#include "itn_internal.h"

namespace {

struct GaugeEdge {
  int src, dst, e;
};

void gauge_level(itn_net* net, const std::vector<GaugeEdge>& es) {
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  const bool cplx = net->cplx;
  const double eps = 2.220446049250313e-16;
  // Gram matrices of the sources
  std::vector<size_t> off(es.size() + 1, 0);
  for (size_t i = 0; i < es.size(); ++i) {
    const size_t c = (size_t)net->edim[es[i].e];
    off[i + 1] = off[i] + c * c * P;
  }
  DevBuf gram(ctx, off.back() * sizeof(double)), rf(ctx, off.back() * sizeof(double)), ri(ctx, off.back() * sizeof(double));
  std::vector<JobSpec> specs(es.size());
  for (size_t i = 0; i < es.size(); ++i) {
    JobSpec& sp = specs[i];
    sp.v = es[i].src;
    sp.open_mask = 1u << (net->slot(es[i].src, es[i].e) + 1);
    sp.out = gram.as<double>() + off[i];
    sp.no_messages = true;
  }
  itn_run_vertex_jobs(net, specs);
  // out[b + chi b'] = sum_x A[x, b] conj(A[x, b']) = conj(G): S = f(out) = conj(f(G)) = f(G)^T
  std::map<int, std::vector<size_t>> by_chi;
  for (size_t i = 0; i < es.size(); ++i) by_chi[net->edim[es[i].e]].push_back(i);
  for (auto& kv : by_chi) {
    std::vector<const double*> in;
    std::vector<double*> o0, o1;
    for (size_t i : kv.second) {
      in.push_back(gram.as<double>() + off[i]);
      o0.push_back(rf.as<double>() + off[i]);
      o1.push_back(ri.as<double>() + off[i]);
    }
    itn_dev_map_eigvals(ctx, cplx, 0, kv.first, (int)in.size(), in.data(), o0.data(), 10.0 * eps);
    itn_dev_map_eigvals(ctx, cplx, 1, kv.first, (int)in.size(), in.data(), o1.data(), 10.0 * eps);
  }
  // Q = A x_bond G^(-1/2): out[.., b, ..] = sum_a A[.., a, ..] G^(-1/2)[a, b] = sum_a A[a] S_inv[b, a]   (trans)
  // B' = R x_bond B:       out[.., b, ..] = sum_a R[b, a] B[.., a, ..]        = sum_a B[a] S[a, b]
  std::map<int, ModeProdSpec> spec_of;  // vertex -> chain (a destination may receive several factors in one level)
  auto spec_for = [&](int v) -> ModeProdSpec& {
    auto it = spec_of.find(v);
    if (it != spec_of.end()) return it->second;
    ModeProdSpec sp;
    memset(&sp, 0, sizeof(sp));
    sp.src = net->T[v].p;
    sp.n = net->T[v].n;
    sp.nm = (int)net->inc[v].size() + 1;
    sp.dims[0] = net->sdim[v];
    for (size_t j = 0; j < net->inc[v].size(); ++j) sp.dims[j + 1] = net->edim[net->inc[v][j]];
    return spec_of.emplace(v, sp).first->second;
  };
  for (size_t i = 0; i < es.size(); ++i) {
    ModeProdSpec& a = spec_for(es[i].src);
    a.mode[a.nsteps] = net->slot(es[i].src, es[i].e) + 1;
    a.mat[a.nsteps] = ri.as<double>() + off[i];
    a.trans[a.nsteps] = 1;
    a.nsteps++;
    ModeProdSpec& b = spec_for(es[i].dst);
    ITN_REQUIRE(b.nsteps < ITN_MAX_MODES, ITN_EUNSUPPORTED, "gauge_walk: too many factors into one vertex");
    b.mode[b.nsteps] = net->slot(es[i].dst, es[i].e) + 1;
    b.mat[b.nsteps] = rf.as<double>() + off[i];
    b.trans[b.nsteps] = 0;
    b.nsteps++;
  }
  std::vector<ModeProdSpec> mps;
  std::vector<int> mv;
  std::vector<std::unique_ptr<DevBuf>> scratch;
  for (auto& kv : spec_of) {
    ModeProdSpec sp = kv.second;
    scratch.emplace_back(new DevBuf(ctx, (size_t)sp.n * P * sizeof(double)));
    sp.w0 = scratch.back()->as<double>();
    if (sp.nsteps > 1) {
      scratch.emplace_back(new DevBuf(ctx, (size_t)sp.n * P * sizeof(double)));
      sp.w1 = scratch.back()->as<double>();
    }
    mps.push_back(sp);
    mv.push_back(kv.first);
  }
  std::vector<const double*> res;
  itn_run_modeprods(ctx, cplx, mps, res);
  for (size_t i = 0; i < mps.size(); ++i) {
    const int v = mv[i];
    CUDA_CHECK(cudaMemcpyAsync(net->T[v].p, res[i], (size_t)mps[i].n * P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    net->touch(v);
  }
}

}  // namespace

#define API_BEGIN try {
#define API_END                              \
  }                                          \
  catch (const ItnError& e) {                \
    itn_set_error(e.what());                 \
    return e.code;                           \
  }                                          \
  catch (const std::exception& e) {          \
    itn_set_error(e.what());                 \
    return ITN_EINVAL;                       \
  }                                          \
  return ITN_OK;

extern "C" int itn_gauge_walk(itn_net* net, const int32_t* src, const int32_t* dst, int n) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  ITN_REQUIRE(n >= 0 && (n == 0 || (src && dst)), ITN_EINVAL, "bad edge sequence");
  ITN_REQUIRE(!net->has_bra(), ITN_EUNSUPPORTED, "gauge_walk acts on a state (no explicit bra layer)");
  ITN_REQUIRE(net->ctx->nranks == 1, ITN_EUNSUPPORTED, "gauge_walk is a sequential walk over a tree: it does not shard");
  CUDA_CHECK(cudaSetDevice(net->ctx->device));
  itn_flush_pending(net);
  std::vector<GaugeEdge> es(n);
  for (int i = 0; i < n; ++i) {
    const int d = net->did(src[i], dst[i]);
    // has_edge check of qr!(tn, edge): "Edge not in graph." (abstractitensornetwork.jl:438 for left_orth!, same for qr!)
    ITN_REQUIRE(d >= 0, ITN_EINVAL, "Edge not in graph.");
    ITN_REQUIRE(net->T[src[i]].p && net->T[dst[i]].p, ITN_EINVAL, "site tensor is not set");
    ITN_REQUIRE(net->edim[d / 2] <= 256, ITN_EUNSUPPORTED, "gauge_walk supports bond extents up to 256");
    es[i] = {src[i], dst[i], d / 2};
  }
  // level of an edge = 1 + the last level that wrote either of its tensors (a source must be final before its Gram matrix
  // is taken; two edges into the same destination at the same level act on different bonds and chain in one launch)
  std::vector<int> wrote(net->nv, -1), srcl(net->nv, -1), level(n, 0);
  int nlev = 0;
  for (int i = 0; i < n; ++i) {
    int lv = std::max(wrote[es[i].src] + 1, srcl[es[i].dst] + 1);
    lv = std::max(lv, srcl[es[i].src] + 1);  // a vertex is the source of at most one edge per level
    // the destination must not be read as a source in this level after this edge rewrote it: sources of a level are read
    // first, all writes follow, so only "written before as destination, now source" needs a new level (wrote[] above)
    level[i] = lv;
    wrote[es[i].src] = std::max(wrote[es[i].src], lv);
    wrote[es[i].dst] = std::max(wrote[es[i].dst], lv);
    srcl[es[i].src] = lv;
    nlev = std::max(nlev, lv + 1);
  }
  for (int l = 0; l < nlev; ++l) {
    std::vector<GaugeEdge> batch;
    for (int i = 0; i < n; ++i)
      if (level[i] == l) batch.push_back(es[i]);
    if (batch.empty()) continue;
    gauge_level(net, batch);  // Q R from the Gram matrix ...
    gauge_level(net, batch);  // ... and once more on Q: orthonormal to eps (CholeskyQR2)
  }
  API_END
}
