// bvh_quality.cpp — offline experiment (CPU): how many node visits / triangle tests per ray do different
// BVH2 builders need on a scene?  LBVH (Morton + highest-differing-bit split = the tree Karras' algorithm
// builds), binned SAH top-down, and PLOC (Meister & Bittner 2018).  Used to decide whether a better GPU
// builder is worth writing.  Input: RTSC file.  Not part of the product.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <random>
#include <vector>
struct V { float x, y, z; };
static V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V operator*(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Box { V lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
  void grow(V p) { lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)}; hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)}; }
  void grow(const Box& b) { grow(b.lo); grow(b.hi); }
  float area() const { V d = hi - lo; return 2 * (d.x * d.y + d.y * d.z + d.z * d.x); } };
struct Tri { V a, b, c; };
struct Node { Box box; int left = -1, right = -1, first = 0, count = 0; };
struct Bvh { std::vector<Node> nodes; std::vector<int> order; int root = 0; };
static std::vector<Tri> tris; static std::vector<Box> tb; static std::vector<V> cen;

static uint64_t spread(uint32_t v) { uint64_t x = v & 0x1fffff; x = (x | x << 32) & 0x1f00000000ffffull; x = (x | x << 16) & 0x1f0000ff0000ffull; x = (x | x << 8) & 0x100f00f00f00f00full; x = (x | x << 4) & 0x10c30c30c30c30c3ull; x = (x | x << 2) & 0x1249249249249249ull; return x; }
static std::vector<uint64_t> mortonSorted(std::vector<int>& order) {
  Box cb; for (auto& c : cen) cb.grow(c);
  int n = (int)tris.size(); std::vector<uint64_t> key(n);
  for (int i = 0; i < n; i++) { auto q = [&](float v, float lo, float hi) { float t = hi > lo ? (v - lo) / (hi - lo) : 0; return (uint32_t)std::min(2097151.0f, t * 2097152.0f); };
    key[i] = spread(q(cen[i].x, cb.lo.x, cb.hi.x)) << 2 | spread(q(cen[i].y, cb.lo.y, cb.hi.y)) << 1 | spread(q(cen[i].z, cb.lo.z, cb.hi.z)); }
  order.resize(n); for (int i = 0; i < n; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<uint64_t> k(n); for (int i = 0; i < n; i++) k[i] = key[order[i]]; return k; }

static Bvh buildLBVH() { Bvh b; auto k = mortonSorted(b.order); int n = (int)tris.size();
  std::function<int(int, int)> rec = [&](int lo, int hi) -> int { int id = (int)b.nodes.size(); b.nodes.push_back(Node());
    if (lo == hi) { b.nodes[id].first = lo; b.nodes[id].count = 1; b.nodes[id].box = tb[b.order[lo]]; return id; }
    int split;
    if (k[lo] == k[hi]) split = (lo + hi) / 2; else { int pre = __builtin_clzll(k[lo] ^ k[hi]); split = lo; int step = hi - lo;
      do { step = (step + 1) >> 1; int ns = split + step; if (ns < hi && __builtin_clzll(k[lo] ^ k[ns]) > pre) split = ns; } while (step > 1); }
    int l = rec(lo, split), r = rec(split + 1, hi); b.nodes[id].left = l; b.nodes[id].right = r; b.nodes[id].box = b.nodes[l].box; b.nodes[id].box.grow(b.nodes[r].box); return id; };
  b.root = rec(0, n - 1); return b; }

static Bvh buildSAH() { Bvh b; int n = (int)tris.size(); b.order.resize(n); for (int i = 0; i < n; i++) b.order[i] = i;
  std::function<int(int, int)> rec = [&](int lo, int hi) -> int { int id = (int)b.nodes.size(); b.nodes.push_back(Node()); Box bb, cb;
    for (int i = lo; i <= hi; i++) { bb.grow(tb[b.order[i]]); cb.grow(cen[b.order[i]]); } b.nodes[id].box = bb;
    if (lo == hi) { b.nodes[id].first = lo; b.nodes[id].count = 1; return id; }
    const int NB = 16; float best = 1e30f; int bax = -1, bsp = -1; V ext = cb.hi - cb.lo;
    for (int ax = 0; ax < 3; ax++) { float e = ax == 0 ? ext.x : ax == 1 ? ext.y : ext.z; if (e <= 0) continue; float lo0 = ax == 0 ? cb.lo.x : ax == 1 ? cb.lo.y : cb.lo.z;
      Box bin[NB]; int cnt[NB] = {0}; for (int i = lo; i <= hi; i++) { V c = cen[b.order[i]]; float v = ax == 0 ? c.x : ax == 1 ? c.y : c.z; int k = std::min(NB - 1, (int)((v - lo0) / e * NB)); bin[k].grow(tb[b.order[i]]); cnt[k]++; }
      float ra[NB]; Box acc; int c = 0; int rc[NB]; for (int k = NB - 1; k > 0; k--) { acc.grow(bin[k]); c += cnt[k]; ra[k] = acc.area(); rc[k] = c; }
      Box l; int lc = 0; for (int k = 0; k < NB - 1; k++) { l.grow(bin[k]); lc += cnt[k]; if (lc == 0 || rc[k + 1] == 0) continue; float cost = l.area() * lc + ra[k + 1] * rc[k + 1]; if (cost < best) { best = cost; bax = ax; bsp = k; } } }
    int mid;
    if (bax < 0) mid = (lo + hi) / 2; else { float e = bax == 0 ? ext.x : bax == 1 ? ext.y : ext.z, lo0 = bax == 0 ? cb.lo.x : bax == 1 ? cb.lo.y : cb.lo.z;
      auto it = std::partition(b.order.begin() + lo, b.order.begin() + hi + 1, [&](int t) { V c = cen[t]; float v = bax == 0 ? c.x : bax == 1 ? c.y : c.z; return std::min(NB - 1, (int)((v - lo0) / e * NB)) <= bsp; });
      mid = (int)(it - b.order.begin()) - 1; if (mid < lo || mid >= hi) mid = (lo + hi) / 2; }
    int l = rec(lo, mid), r = rec(mid + 1, hi); b.nodes[id].left = l; b.nodes[id].right = r; return id; };
  b.root = rec(0, n - 1); return b; }

static Bvh buildPLOC(int radius) { Bvh b; mortonSorted(b.order); int n = (int)tris.size();
  b.nodes.resize(n); std::vector<int> cl(n); for (int i = 0; i < n; i++) { b.nodes[i].box = tb[b.order[i]]; b.nodes[i].first = i; b.nodes[i].count = 1; cl[i] = i; }
  int iters = 0;
  while (cl.size() > 1) { int m = (int)cl.size(); std::vector<int> nn(m);
    for (int i = 0; i < m; i++) { float best = 1e30f; int bj = -1; for (int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); j++) { if (j == i) continue; Box u = b.nodes[cl[i]].box; u.grow(b.nodes[cl[j]].box); float a = u.area(); if (a < best) { best = a; bj = j; } } nn[i] = bj; }
    std::vector<int> next; for (int i = 0; i < m; i++) { int j = nn[i]; if (nn[j] == i) { if (i < j) { Node nd; nd.left = cl[i]; nd.right = cl[j]; nd.box = b.nodes[cl[i]].box; nd.box.grow(b.nodes[cl[j]].box); b.nodes.push_back(nd); next.push_back((int)b.nodes.size() - 1); } } else next.push_back(cl[i]); }
    cl.swap(next); iters++; }
  b.root = cl[0]; fprintf(stderr, "PLOC r=%d iterations %d\n", radius, iters); return b; }

// Kensler-style tree rotations: for every inner node try swapping a child with a grandchild on the other side;
// keep the swap that lowers the summed surface area most.  `passes` bottom-up sweeps.
static void rotate(Bvh& b, int passes) {
  std::vector<int> order; std::function<void(int)> post = [&](int n) { if (b.nodes[n].count) return; post(b.nodes[n].left); post(b.nodes[n].right); order.push_back(n); };
  for (int p = 0; p < passes; p++) { order.clear(); post(b.root); int applied = 0;
    for (int n : order) { Node& N = b.nodes[n]; int best = -1; float gain = 0;
      for (int side = 0; side < 2; side++) { int c = side ? N.right : N.left, o = side ? N.left : N.right; if (b.nodes[o].count) continue; // swap c with a child of o
        for (int k = 0; k < 2; k++) { int g = k ? b.nodes[o].right : b.nodes[o].left, keep = k ? b.nodes[o].left : b.nodes[o].right; Box nb = b.nodes[keep].box; nb.grow(b.nodes[c].box); float ga = b.nodes[o].box.area() - nb.area(); if (ga > gain) { gain = ga; best = side * 2 + k; } (void)g; } }
      if (best >= 0) { int side = best >> 1, k = best & 1; int& c = side ? N.right : N.left; int o = side ? N.left : N.right; int& g = k ? b.nodes[o].right : b.nodes[o].left; std::swap(c, g); int keep = k ? b.nodes[o].left : b.nodes[o].right; Box nb = b.nodes[keep].box; nb.grow(b.nodes[g].box); b.nodes[o].box = nb; applied++; } }
    fprintf(stderr, "rotation pass %d applied %d\n", p, applied); } }

static double sahCost(const Bvh& b) { double c = 0; double ra = b.nodes[b.root].box.area(); for (auto& n : b.nodes) c += n.box.area() / ra * (n.count ? 1.0 : 1.2); return c; }
static int depthOf(const Bvh& b, int n) { return b.nodes[n].count ? 1 : 1 + std::max(depthOf(b, b.nodes[n].left), depthOf(b, b.nodes[n].right)); }

static bool slab(const Box& bx, V o, V inv, float tmax, float& tn) { float t0 = (bx.lo.x - o.x) * inv.x, t1 = (bx.hi.x - o.x) * inv.x; float a = std::min(t0, t1), z = std::max(t0, t1);
  t0 = (bx.lo.y - o.y) * inv.y; t1 = (bx.hi.y - o.y) * inv.y; a = std::max(a, std::min(t0, t1)); z = std::min(z, std::max(t0, t1));
  t0 = (bx.lo.z - o.z) * inv.z; t1 = (bx.hi.z - o.z) * inv.z; a = std::max(a, std::min(t0, t1)); z = std::min(z, std::max(t0, t1)); tn = std::max(a, 0.0f); return tn <= std::min(z, tmax); }
static bool hitTri(const Tri& t, V o, V d, float& dst) { V e0 = t.b - t.a, e1 = t.c - t.a, N = cross(e0, e1); float det = -dot(d, N); if (det < 1e-10f) return false; float inv = 1 / det; V ao = o - t.a; dst = dot(ao, N) * inv; if (dst <= 1e-6f) return false; V dao = cross(d, ao); float u = -dot(e1, dao) * inv, v = dot(e0, dao) * inv; return u >= 0 && v >= 0 && 1 - u - v >= 0; }
struct Stats { double nodes = 0, tris = 0; };
static float trace(const Bvh& b, V o, V d, Stats& st, int& hitTriIdx) { V inv{1 / d.x, 1 / d.y, 1 / d.z}; float best = 1e30f; hitTriIdx = -1; int stack[128]; float ts[128]; int sp = 0; int cur = b.root;
  for (;;) { const Node& n = b.nodes[cur];
    if (n.count) { for (int i = n.first; i < n.first + n.count; i++) { st.tris++; float dst; if (hitTri(tris[b.order[i]], o, d, dst) && dst < best) { best = dst; hitTriIdx = b.order[i]; } } }
    else { st.nodes++; float tl, tr; bool hl = slab(b.nodes[n.left].box, o, inv, best, tl), hr = slab(b.nodes[n.right].box, o, inv, best, tr);
      if (hl && hr) { bool lf = tl <= tr; stack[sp] = lf ? n.right : n.left; ts[sp++] = lf ? tr : tl; cur = lf ? n.left : n.right; continue; } if (hl) { cur = n.left; continue; } if (hr) { cur = n.right; continue; } }
    bool got = false; while (sp > 0) { --sp; if (ts[sp] <= best) { cur = stack[sp]; got = true; break; } } if (!got) break; }
  return best; }


// ---- wide BVH (round-2 costing): collapse a BVH2 into k-wide nodes by repeatedly opening the child with the
// largest surface area; ordered traversal (children sorted by entry distance, far ones pushed).
struct WNode { int n = 0; int child[8]; Box box[8]; };  // child >= 0: wide node, < 0: ~leaf-node index of the BVH2
struct WBvh { std::vector<WNode> nodes; };
static WBvh collapse(const Bvh& b, int width) { WBvh w; w.nodes.reserve(b.nodes.size() / 2);
  std::function<int(int)> rec = [&](int root2) -> int { int id = (int)w.nodes.size(); w.nodes.push_back(WNode()); std::vector<int> slots{b.nodes[root2].left, b.nodes[root2].right};
    for (;;) { if ((int)slots.size() >= width) break; int best = -1; float ba = -1; for (size_t i = 0; i < slots.size(); i++) if (!b.nodes[slots[i]].count) { float a = b.nodes[slots[i]].box.area(); if (a > ba) { ba = a; best = (int)i; } }
      if (best < 0) break; int open = slots[best]; slots[best] = b.nodes[open].left; slots.push_back(b.nodes[open].right); }
    WNode nd; nd.n = (int)slots.size(); for (int i = 0; i < nd.n; i++) { nd.box[i] = b.nodes[slots[i]].box; nd.child[i] = b.nodes[slots[i]].count ? ~slots[i] : 0; }
    w.nodes[id] = nd; for (int i = 0; i < nd.n; i++) if (!b.nodes[slots[i]].count) { int c = rec(slots[i]); w.nodes[id].child[i] = c; } return id; };
  rec(b.root); return w; }
struct WStats { double visits = 0, boxes = 0, tris = 0, pushes = 0; };
static void traceWide(const Bvh& b, const WBvh& w, V o, V d, WStats& st) { V inv{1 / d.x, 1 / d.y, 1 / d.z}; float best = 1e30f; int stack[256]; float ts[256]; int sp = 0; int cur = 0;
  for (;;) { if (cur >= 0) { const WNode& n = w.nodes[cur]; st.visits++; st.boxes += n.n; int idx[8]; float tn[8]; int h = 0;
      for (int i = 0; i < n.n; i++) { float t; if (slab(n.box[i], o, inv, best, t)) { int k = h++; while (k > 0 && tn[k - 1] > t) { tn[k] = tn[k - 1]; idx[k] = idx[k - 1]; k--; } tn[k] = t; idx[k] = i; } }
      for (int k = h - 1; k >= 1; k--) { stack[sp] = n.child[idx[k]]; ts[sp++] = tn[k]; st.pushes++; }
      if (h) { cur = n.child[idx[0]]; continue; } }
    else { const Node& l = b.nodes[~cur]; for (int i = l.first; i < l.first + l.count; i++) { st.tris++; float dst; if (hitTri(tris[b.order[i]], o, d, dst) && dst < best) best = dst; } }
    bool got = false; while (sp > 0) { --sp; if (ts[sp] <= best) { cur = stack[sp]; got = true; break; } } if (!got) break; } }

int main(int argc, char** argv) { if (argc < 2) return 2; FILE* f = fopen(argv[1], "rb"); int64_t hdr[8]; if (!f || fread(hdr, 8, 8, f) != 8) return 1; int n = (int)hdr[1]; std::vector<float> raw((size_t)n * 20); if (fread(raw.data(), 80, n, f) != (size_t)n) return 1; fclose(f);
  tris.resize(n); tb.resize(n); cen.resize(n); for (int i = 0; i < n; i++) { float* p = &raw[(size_t)i * 20]; tris[i] = {{p[0], p[1], p[2]}, {p[4], p[5], p[6]}, {p[8], p[9], p[10]}}; tb[i].grow(tris[i].a); tb[i].grow(tris[i].b); tb[i].grow(tris[i].c); cen[i] = (tb[i].lo + tb[i].hi) * 0.5f; }
  // rays: diffuse bounces — start on a random triangle (area-agnostic: via a primary hit from a random interior point), cosine-ish direction
  std::mt19937 rng(7); std::uniform_real_distribution<float> U(0, 1); Box sb; for (auto& b : tb) sb.grow(b);
  Bvh ref = buildSAH(); std::vector<V> ro, rd; Stats dummy;
  while ((int)ro.size() < 200000) { V o{sb.lo.x + (sb.hi.x - sb.lo.x) * U(rng), sb.lo.y + (sb.hi.y - sb.lo.y) * U(rng), sb.lo.z + (sb.hi.z - sb.lo.z) * U(rng)}; V d{U(rng) * 2 - 1, U(rng) * 2 - 1, U(rng) * 2 - 1}; float l = std::sqrt(dot(d, d)); if (l > 1 || l < 1e-3f) continue; d = d * (1 / l);
    int ti; float t = trace(ref, o, d, dummy, ti); if (ti < 0) continue; V p = o + d * (t * 1.001f); const Tri& T = tris[ti]; V N = cross(T.b - T.a, T.c - T.a); N = N * (1 / std::sqrt(dot(N, N)));
    V r; do { r = {U(rng) * 2 - 1, U(rng) * 2 - 1, U(rng) * 2 - 1}; } while (dot(r, r) >= 1); r = r * (1 / std::sqrt(dot(r, r))); V nd = N + r; float nl = std::sqrt(dot(nd, nd)); if (nl < 1e-4f) continue; ro.push_back(p); rd.push_back(nd * (1 / nl)); }
  auto eval = [&](const char* name, const Bvh& b) { Stats st; int ti; for (size_t i = 0; i < ro.size(); i++) trace(b, ro[i], rd[i], st, ti); printf("%-12s nodes %8zu depth %3d SAH %8.2f  node visits/ray %7.2f  tri tests/ray %6.2f\n", name, b.nodes.size(), depthOf(b, b.root), sahCost(b), st.nodes / ro.size(), st.tris / ro.size()); fflush(stdout); };
  eval("binned SAH", ref); { Bvh l = buildLBVH(); eval("LBVH", l); } for (int r : {10, 25}) { char nm[32]; snprintf(nm, 32, "PLOC r=%d", r); Bvh p = buildPLOC(r); eval(nm, p); if (r == 10) { for (int k = 0; k < 3; k++) { rotate(p, 2); snprintf(nm, 32, "PLOC+rot%d", 2 * (k + 1)); eval(nm, p); } } }
  { Bvh p = buildPLOC(10); for (int width : {2, 4, 8}) { WBvh w = collapse(p, width); WStats st; for (size_t i = 0; i < ro.size(); i++) traceWide(p, w, ro[i], rd[i], st); double m = (double)ro.size();
      printf("PLOC -> BVH%d  wide nodes %8zu  visits/ray %6.2f  box tests/ray %6.2f  pushes/ray %5.2f  tri tests/ray %5.2f\n", width, w.nodes.size(), st.visits / m, st.boxes / m, st.pushes / m, st.tris / m); fflush(stdout); } }
  return 0; }
