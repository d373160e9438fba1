// pf_bme.h -- host-only tree building from a distance matrix: BIONJ start tree, then balanced
// minimum evolution (BME) hill climbing by NNI and by SPR.  No device work.
//
// This is what the reference's README does with the matrices this path produces (README.md:85-92:
// `fastme -i x.phy -o x.nwk --nni --spr`, FastME 2.1.6.4 shipped as bin/bin_linux/fastme): build a BIONJ
// tree (FastME's default start method), run one search with balanced NNIs and one with balanced SPRs from
// it, keep the shorter tree, branch lengths by the balanced formulas.  FastME's source is not part of the
// reference tree (only the binary is); the algorithms are restated from the publications:
//   Gascuel 1997 (BIONJ), Desper & Gascuel 2002 (balanced minimum evolution, Pauplin's tree length,
//   balanced NNI), Hordijk & Gascuel 2005 (SPR under BME), Lefort, Desper & Gascuel 2015 (FastME 2.0).
// tests/test_bme_cpu.py pins the result on trees written by the FastME binary itself for the 20 reference
// matrices (tests/golden/ref_trees_pf.json) and on a brute-force evaluation of Pauplin's formula.
//
// Notation: a directed edge h = (v -> w) stands for the subtree hanging on w's side.  avg(h1, h2) is the
// balanced average distance between two disjoint subtrees: d_ij for two leaves, otherwise the mean of the
// two child subtrees' averages (every split halves the weight, whatever the subtree sizes).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace pfbme {

struct Tree {
  int n = 0;                                // leaves 0..n-1, internal nodes n..2n-3
  std::vector<std::array<int, 3>> adj;      // node -> incident edge ids (-1: unused slot of a leaf)
  std::vector<std::array<int, 2>> ends;     // edge -> (u, v); directed edge 2e = u->v, 2e+1 = v->u
  int n_nodes() const { return (int)adj.size(); }
  int n_edges() const { return (int)ends.size(); }
  int head(int h) const { return ends[h >> 1][(h & 1) ^ 1]; }
  int tail(int h) const { return ends[h >> 1][h & 1]; }
  int dir(int from, int e) const { return 2 * e + (ends[e][0] == from ? 0 : 1); }   // directed edge leaving `from` along e
  // the two directed edges that continue h beyond its head (head must be internal)
  void children(int h, int& c1, int& c2) const {
    const int w = head(h), e = h >> 1;
    int out[2], k = 0;
    for (int s = 0; s < 3; ++s)
      if (adj[w][s] != e) out[k++] = dir(w, adj[w][s]);
    c1 = out[0]; c2 = out[1];
  }
  void replace_edge(int node, int old_e, int new_e) {
    for (int s = 0; s < 3; ++s)
      if (adj[node][s] == old_e) { adj[node][s] = new_e; return; }
  }
  void replace_end(int e, int old_node, int new_node) {
    if (ends[e][0] == old_node) ends[e][0] = new_node; else ends[e][1] = new_node;
  }
};

// Balanced averages between all pairs of disjoint subtrees of the current topology: one H x H table
// (H directed edges), rebuilt after every accepted move.  Directed edges are visited by increasing subtree
// size, so a row is the mean of its two child rows (one contiguous pass per row); rows of leaf-pointing edges
// are filled along the same order from the distance matrix.  Entries of overlapping subtrees are meaningless
// and never read.
struct Averages {
  const Tree& t;
  const double* D;
  int H;
  std::vector<double> val;
  std::vector<int> c1, c2, order;
  Averages(const Tree& tree, const double* dist) : t(tree), D(dist), H(2 * tree.n_edges()) {
    val.resize((size_t)H * H);
    c1.resize((size_t)H); c2.resize((size_t)H); order.resize((size_t)H);
    fill();
  }
  void reset() { fill(); }
  double get(int a, int b) const { return val[(size_t)a * H + b]; }
  void fill() {
    if (H == 0) return;
    const int n = t.n;
    // subtree sizes: depth-first from leaf 0; (parent -> v) holds the leaves below v, (v -> parent) the rest
    std::vector<int> size((size_t)H, 0), stack, parent_edge((size_t)t.n_nodes(), -1), seq;
    stack.push_back(0);
    while (!stack.empty()) {
      const int v = stack.back();
      stack.pop_back();
      seq.push_back(v);
      for (int k = 0; k < 3; ++k) {
        const int e = t.adj[v][k];
        if (e < 0 || e == parent_edge[v]) continue;
        const int w = (t.ends[e][0] == v) ? t.ends[e][1] : t.ends[e][0];
        parent_edge[w] = e;
        stack.push_back(w);
      }
    }
    std::vector<int> below((size_t)t.n_nodes(), 0);
    for (size_t q = seq.size(); q-- > 1;) {       // children before parents
      const int v = seq[q], e = parent_edge[v];
      if (v < n) below[v] = 1;
      const int par = (t.ends[e][0] == v) ? t.ends[e][1] : t.ends[e][0];
      below[par] += below[v];
      size[t.dir(par, e)] = below[v];
      size[t.dir(v, e)] = n - below[v];
    }
    for (int h = 0; h < H; ++h) {
      order[h] = h;
      if (t.head(h) >= n) t.children(h, c1[h], c2[h]); else c1[h] = c2[h] = -1;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return size[a] < size[b]; });
    for (int a : order) {
      double* row = &val[(size_t)a * H];
      if (c1[a] < 0) {
        const double* drow = D + (size_t)t.head(a) * n;
        for (int b : order) row[b] = (c1[b] < 0) ? drow[t.head(b)] : 0.5 * (row[c1[b]] + row[c2[b]]);
      } else {
        const double* r1 = &val[(size_t)c1[a] * H];
        const double* r2 = &val[(size_t)c2[a] * H];
        for (int b = 0; b < H; ++b) row[b] = 0.5 * (r1[b] + r2[b]);
      }
    }
  }
};

// Balanced branch length of edge e (Desper & Gascuel 2002, eq. for internal and external edges).
inline double edge_length(const Tree& t, Averages& A, int e) {
  const int u = t.ends[e][0], v = t.ends[e][1];
  if (u >= t.n && v >= t.n) {
    int a, b, c, d;
    t.children(2 * e + 1, a, b);   // subtrees on u's side
    t.children(2 * e, c, d);       // subtrees on v's side
    return 0.25 * (A.get(a, c) + A.get(a, d) + A.get(b, c) + A.get(b, d)) - 0.5 * (A.get(a, b) + A.get(c, d));
  }
  const int leaf_dir = (v < t.n) ? 2 * e : 2 * e + 1;   // directed edge pointing at the leaf
  if (t.head(leaf_dir ^ 1) < t.n) return A.get(leaf_dir, leaf_dir ^ 1);   // two-leaf tree
  int a, b;
  t.children(leaf_dir ^ 1, a, b);
  return 0.5 * (A.get(leaf_dir, a) + A.get(leaf_dir, b) - A.get(a, b));
}

inline double tree_length(const Tree& t, Averages& A) {
  double s = 0.0;
  for (int e = 0; e < t.n_edges(); ++e) s += edge_length(t, A, e);
  return s;
}

// ---- BIONJ (Gascuel 1997) ----------------------------------------------------------------------
// Agglomeration with the NJ criterion Q_ij = (r-2) d_ij - S_i - S_j and the variance-weighted
// reduction d_uk = lambda (d_ik - l_i) + (1 - lambda)(d_jk - l_j); the first minimum in the scan order
// x = 0..n-1, y < x wins and a later pair must beat it by more than 1e-6 (the published implementation's
// tie rule).  The new cluster takes the slot of x.  With bionj = false lambda is 1/2 (plain NJ).
inline Tree bionj(const double* D, int n, bool use_bionj = true, std::vector<double>* own_lengths = nullptr) {
  Tree t;
  std::vector<double> blen;   // the agglomeration's own branch lengths, per edge
  t.n = n;
  if (n < 2) return t;      // a single taxon: no edges
  t.adj.assign((size_t)(2 * n - 2), {-1, -1, -1});
  t.ends.reserve((size_t)(2 * n - 3));
  auto connect = [&](int a, int b) {
    const int e = (int)t.ends.size();
    t.ends.push_back({a, b});
    for (int node : {a, b})
      for (int s = 0; s < 3; ++s)
        if (t.adj[node][s] < 0) { t.adj[node][s] = e; break; }
  };
  if (n == 2) {
    connect(0, 1);
    if (own_lengths) own_lengths->assign(1, D[1]);
    return t;
  }
  std::vector<double> d((size_t)n * n), var((size_t)n * n), S((size_t)n);
  for (size_t i = 0; i < (size_t)n * n; ++i) d[i] = var[i] = D[i];
  std::vector<int> node((size_t)n);      // tree node currently represented by slot i
  std::vector<char> alive((size_t)n, 1);
  for (int i = 0; i < n; ++i) node[i] = i;
  int next_node = n, r = n;
  auto dd = [&](int i, int j) -> double& { return d[(size_t)i * n + j]; };
  auto vv = [&](int i, int j) -> double& { return var[(size_t)i * n + j]; };
  while (r > 3) {
    for (int i = 0; i < n; ++i) {
      if (!alive[i]) continue;
      double s = 0.0;
      for (int j = 0; j < n; ++j)
        if (alive[j] && j != i) s += dd(i, j);
      S[i] = s;
    }
    double qmin = 1.0e300;
    int a = -1, b = -1;
    for (int x = 0; x < n; ++x) {
      if (!alive[x]) continue;
      for (int y = 0; y < x; ++y) {
        if (!alive[y]) continue;
        const double q = (r - 2) * dd(x, y) - S[x] - S[y];
        if (q < qmin - 0.000001) { qmin = q; a = x; b = y; }
      }
    }
    const double dab = dd(a, b), vab = vv(a, b);
    const double la = 0.5 * (dab + (S[a] - S[b]) / (r - 2));
    const double lb = 0.5 * (dab + (S[b] - S[a]) / (r - 2));
    double lambda = 0.5;
    if (use_bionj && vab != 0.0) {
      double acc = 0.0;
      for (int i = 0; i < n; ++i)
        if (alive[i] && i != a && i != b) acc += vv(b, i) - vv(a, i);
      lambda = 0.5 + acc / (2.0 * (r - 2) * vab);
      if (lambda > 1.0) lambda = 1.0;
      if (lambda < 0.0) lambda = 0.0;
    }
    for (int i = 0; i < n; ++i) {
      if (!alive[i] || i == a || i == b) continue;
      const double du = lambda * (dd(a, i) - la) + (1.0 - lambda) * (dd(b, i) - lb);
      const double vu = lambda * vv(a, i) + (1.0 - lambda) * vv(b, i) - lambda * (1.0 - lambda) * vab;
      dd(a, i) = dd(i, a) = du;
      vv(a, i) = vv(i, a) = vu;
    }
    const int u = next_node++;
    connect(u, node[a]); blen.push_back(la);
    connect(u, node[b]); blen.push_back(lb);
    node[a] = u;
    alive[b] = 0;
    --r;
  }
  const int u = next_node++;
  int last[3], k = 0;
  for (int i = 0; i < n; ++i)
    if (alive[i]) last[k++] = i;
  for (int m = 0; m < 3; ++m) {
    const int i = last[m], j = last[(m + 1) % 3], l = last[(m + 2) % 3];
    connect(u, node[i]);
    blen.push_back(0.5 * (dd(i, j) + dd(i, l) - dd(j, l)));
  }
  if (own_lengths) *own_lengths = blen;
  return t;
}

// ---- balanced NNI (Desper & Gascuel 2002) ------------------------------------------------------
// Around an internal edge with subtrees (a, b | c, d), swapping b and c changes the balanced length by
// [(avg(a,c) + avg(b,d)) - (avg(a,b) + avg(c,d))] / 4 (likewise b and d).  Best improvement first,
// until no NNI shortens the tree by more than eps.  Returns the number of NNIs performed.
inline int bme_nni(Tree& t, const double* D, double eps, int max_moves = 1 << 30) {
  if (t.n < 4) return 0;
  Averages A(t, D);
  int moves = 0;
  while (moves < max_moves) {
    double best = -eps;
    int be = -1, bx = -1, by = -1;
    for (int e = 0; e < t.n_edges(); ++e) {
      if (t.ends[e][0] < t.n || t.ends[e][1] < t.n) continue;
      int a, b, c, d;
      t.children(2 * e + 1, a, b);
      t.children(2 * e, c, d);
      const double base = A.get(a, b) + A.get(c, d);
      const double d1 = 0.25 * (A.get(a, c) + A.get(b, d) - base);   // swap b <-> c
      const double d2 = 0.25 * (A.get(a, d) + A.get(b, c) - base);   // swap b <-> d
      if (d1 < best) { best = d1; be = e; bx = b; by = c; }
      if (d2 < best) { best = d2; be = e; bx = b; by = d; }
    }
    if (be < 0) break;
    const int u = t.ends[be][0], v = t.ends[be][1];   // bx hangs on u, by on v: exchange the attachments
    const int ex = bx >> 1, ey = by >> 1;
    t.replace_edge(u, ex, ey);
    t.replace_edge(v, ey, ex);
    t.replace_end(ex, u, v);
    t.replace_end(ey, v, u);
    A.reset();
    ++moves;
  }
  return moves;
}

// ---- SPR under BME (Hordijk & Gascuel 2005) ----------------------------------------------------
// Pruning the subtree S of directed edge (p -> q) and regrafting it k edges away equals k successive
// NNIs that carry S across one vertex at a time; each step costs
//   [(avg(behind, B) + avg(S, R)) - (avg(behind, S) + avg(B, R))] / 4
// where B is the side subtree of the crossed vertex, R the rest ahead, and `behind` everything S has
// already passed, a weighted mix of genuine subtrees of the unmodified tree (weights halve per step).
struct SprSearch {
  Tree& t;
  Averages A;
  double best;
  int best_s, best_target;
  SprSearch(Tree& tree, const double* D) : t(tree), A(tree, D), best(0), best_s(-1), best_target(-1) {}
  // S (directed edge s, first neighbour subtree a0) is about to cross head(r) after k earlier crossings.
  // `behind` = a0 joined with the side subtrees passed so far; together with S it is the subtree U of r
  // reversed in the unmodified tree, where S carries the weight a0 has in `behind`:
  //   avg(behind, X) = avg(U, X) - 2^-(k+1) (avg(S, X) - avg(a0, X)),
  // and avg(behind, S) follows the crossings: bs <- (bs + avg(B, S)) / 2.
  void walk(int s, int a0, int r, double cum, double scale, double bs) {
    if (t.head(r) < t.n) return;
    const int z[2] = {A.c1[r], A.c2[r]};
    const double* row_s = &A.val[(size_t)s * A.H];
    const double* row_a = &A.val[(size_t)a0 * A.H];
    const double* row_u = &A.val[(size_t)(r ^ 1) * A.H];
    const double bb = A.get(z[0], z[1]);
    for (int k = 0; k < 2; ++k) {
      const int R = z[k], B = z[k ^ 1];
      const double behind_b = row_u[B] - scale * (row_s[B] - row_a[B]);
      const double delta = cum + 0.25 * (behind_b + row_s[R] - bs - bb);
      if (delta < best) { best = delta; best_s = s; best_target = R >> 1; }
      walk(s, a0, R, delta, 0.5 * scale, 0.5 * (bs + row_s[B]));
    }
  }
  void scan(double eps) {
    best = -eps; best_s = best_target = -1;
    for (int s = 0; s < 2 * t.n_edges(); ++s) {
      const int p = t.tail(s);
      if (p < t.n) continue;                 // a subtree is pruned from an internal vertex
      const int x = A.c1[s ^ 1], y = A.c2[s ^ 1];   // the two other directed edges leaving p
      walk(s, y, x, 0.0, 0.5, A.get(y, s));
      walk(s, x, y, 0.0, 0.5, A.get(x, s));
    }
  }
  void apply() {   // move p (with S attached) into the middle of the target edge
    const int s = best_s, p = t.tail(s), es = s >> 1;
    int ex = -1, ey = -1;
    for (int k = 0; k < 3; ++k)
      if (t.adj[p][k] != es) { if (ex < 0) ex = t.adj[p][k]; else ey = t.adj[p][k]; }
    const int y = (t.ends[ey][0] == p) ? t.ends[ey][1] : t.ends[ey][0];
    // close the gap: ex now runs x - y, ey is free
    t.replace_end(ex, p, y);
    t.replace_edge(y, ey, ex);
    // open the target edge (a, b): it keeps a - p, ey becomes p - b
    const int et = (best_target == ey) ? ex : best_target;   // (cannot happen: ex / ey are never targets)
    const int b = t.ends[et][1];
    t.replace_end(et, b, p);
    t.replace_edge(b, et, ey);
    t.ends[ey] = {p, b};
    t.replace_edge(p, ex, et);
    A.reset();
  }
};

inline int bme_spr(Tree& t, const double* D, double eps, int max_moves = 1 << 30) {
  if (t.n < 5) return 0;   // with four leaves every SPR is an NNI
  SprSearch S(t, D);
  int moves = 0;
  while (moves < max_moves) {
    S.scan(eps);
    if (S.best_s < 0) break;
    S.apply();
    ++moves;
  }
  return moves;
}

// Newick text rooted at the last internal node (trifurcation), balanced branch lengths.
inline std::string newick(const Tree& t, const double* D, const std::vector<std::string>& labels, int digits,
                          bool clip_negative, const std::vector<double>* lengths = nullptr) {
  Averages A(t, D);
  auto edge_len = [&](int e) { return lengths ? (*lengths)[(size_t)e] : edge_length(t, A, e); };
  auto fmt = [&](double x) {
    char b[420];
    if (clip_negative && x < 0) x = 0.0;
    snprintf(b, sizeof(b), "%.*f", digits, x);
    return std::string(b);
  };
  if (t.n == 1) return labels[0] + ";";
  if (t.n == 2) {
    const double h = 0.5 * D[1];
    return "(" + labels[0] + ":" + fmt(h) + "," + labels[1] + ":" + fmt(h) + ");";
  }
  // iterative post-order from the root (explicit stack: caterpillar trees are n deep)
  const int root = t.n_nodes() - 1;
  struct Frame { int node, via, k; std::string text; };
  std::vector<Frame> st;
  st.push_back({root, -1, 0, "("});
  std::string result;
  while (!st.empty()) {
    Frame& f = st.back();
    if (f.node < t.n) {
      std::string leaf = labels[f.node] + ":" + fmt(edge_len(f.via));
      st.pop_back();
      Frame& par = st.back();
      if (par.text.size() > 1) par.text += ",";
      par.text += leaf;
      continue;
    }
    if (f.k < 3) {
      const int e = t.adj[f.node][f.k++];
      if (e == f.via) continue;
      const int child = (t.ends[e][0] == f.node) ? t.ends[e][1] : t.ends[e][0];
      st.push_back({child, e, 0, "("});
      continue;
    }
    std::string done = f.text + ")";
    if (f.via >= 0) done += ":" + fmt(edge_len(f.via));
    st.pop_back();
    if (st.empty()) { result = done + ";"; break; }
    Frame& par = st.back();
    if (par.text.size() > 1) par.text += ",";
    par.text += done;
  }
  return result;
}

struct Result {
  std::string newick;
  double length_own;     // start tree with the agglomeration's own branch lengths
  double length_start;   // start tree, balanced lengths (what the searches start from)
  double length_nni, length_spr;
  int n_nni, n_spr;
  int kept;              // 0 start tree, 1 NNI result, 2 SPR result
};

// flags: bit 0 = NNI search, bit 1 = SPR search, bit 2 = plain NJ start tree instead of BIONJ.
// Without a search the start tree is written with the agglomeration's own branch lengths (`fastme -i x` alone).
// With a search, FastME keeps the shortest of { start tree with its own (BIONJ) branch lengths, NNI result, SPR result };
// observed on the binary: whenever the BIONJ length is below both balanced lengths the output is the BIONJ tree
// with BIONJ branch lengths (tools/bme_vs_fastme.py).
inline Result build(const double* D, int n, const std::vector<std::string>& labels, int flags, int digits = 8,
                    bool clip_negative = false, double eps = 1e-9) {
  Result r{};
  std::vector<double> own;
  Tree start = bionj(D, n, !(flags & 4), &own);
  for (double x : own) r.length_own += x;
  {
    Averages A(start, D);
    r.length_start = (n >= 2) ? tree_length(start, A) : 0.0;
  }
  r.length_nni = r.length_spr = r.length_start;
  if (!(flags & 3) || n < 4) {
    r.newick = newick(start, D, labels, digits, clip_negative, &own);
    return r;
  }
  const Tree* out = &start;
  double best = r.length_own;
  Tree t_nni, t_spr;
  if (flags & 1) {
    t_nni = start;
    r.n_nni = bme_nni(t_nni, D, eps);
    Averages A(t_nni, D);
    r.length_nni = tree_length(t_nni, A);
    // (FastME's BIONJ length carries single-precision noise of ~1e-8; on four taxa, where both lengths are equal in
    //  exact arithmetic, the binary ends on the balanced tree: a tie goes to the search result)
    if (r.length_nni < best + 1e-7) { best = std::min(best, r.length_nni); out = &t_nni; r.kept = 1; }
  }
  if (flags & 2) {
    t_spr = start;
    r.n_spr = bme_spr(t_spr, D, eps);
    Averages A(t_spr, D);
    r.length_spr = tree_length(t_spr, A);
    if (r.length_spr < best + (r.kept == 0 ? 1e-7 : 0.0)) { best = std::min(best, r.length_spr); out = &t_spr; r.kept = 2; }
  }
  r.newick = newick(*out, D, labels, digits, clip_negative, r.kept == 0 ? &own : nullptr);
  return r;
}

}  // namespace pfbme
