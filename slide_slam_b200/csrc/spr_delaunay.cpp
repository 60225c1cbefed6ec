// spr_delaunay.cpp -- 2-D Delaunay triangulation of a landmark map on the host: the triangle source of
// the SlideGraph half.  Replaces Observation::delaunayTriangulation
// (clipper_semantic_object/src/triangulation/observation.cpp:13-88), which calls qhull with
// "Qt Qbb Qc Qz Q12 d"; SURVEY.md section 8 row a9 keeps this stage on the CPU (O(n log n), once per map).
//
// Sweep-hull construction: the points are inserted in order of distance from the circumcentre of a
// seed triangle; each point is joined to the edges of the current convex hull it can see, and the new
// triangles are legalised by edge flips (in-circle test).  Predicates are evaluated in double precision
// with a forward error bound and re-evaluated in 80-bit extended precision when the bound does not
// decide -- exact for the integer-valued coordinates of the reference's own test scene
// (place_recognition_test.cpp:13-28).  A value of exactly zero means "cocircular": no flip, which keeps
// any of the (equally valid) triangulations qhull's "Qt" option could return.
//
// For point sets in general position the Delaunay triangulation is unique, so the triangle SET equals
// qhull's; the ORDER of the triangles and of the three vertices of a triangle is this file's, not
// qhull's facet order (INTEGRATION.md states what that does and does not change downstream).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "spr_delaunay.h"

namespace spr {
namespace {

const double kEps = std::numeric_limits<double>::epsilon();

// > 0: a, b, c in counter-clockwise order
double orient(double ax, double ay, double bx, double by, double cx, double cy) {
  const double l = (ay - cy) * (bx - cx), r = (ax - cx) * (by - cy);
  const double det = r - l;
  const double bound = 4.0 * kEps * (std::fabs(l) + std::fabs(r));
  if (std::fabs(det) > bound) return det;
  const long double L = ((long double)ay - cy) * ((long double)bx - cx), R = ((long double)ax - cx) * ((long double)by - cy);
  return (double)(R - L);
}

// > 0: p strictly inside the circumcircle of the counter-clockwise triangle a, b, c
double in_circle(double ax, double ay, double bx, double by, double cx, double cy, double px, double py) {
  const double dx = ax - px, dy = ay - py, ex = bx - px, ey = by - py, fx = cx - px, fy = cy - py;
  const double ap = dx * dx + dy * dy, bp = ex * ex + ey * ey, cp = fx * fx + fy * fy;
  const double t1 = dx * (ey * cp - bp * fy), t2 = dy * (ex * cp - bp * fx), t3 = ap * (ex * fy - ey * fx);
  const double det = t1 - t2 + t3;
  const double mag = (std::fabs(dx) + std::fabs(dy)) * (std::fabs(ey * cp) + std::fabs(bp * fy) + std::fabs(ex * cp) + std::fabs(bp * fx)) +
                     ap * (std::fabs(ex * fy) + std::fabs(ey * fx));
  if (std::fabs(det) > 16.0 * kEps * mag) return det;
  const long double Dx = (long double)ax - px, Dy = (long double)ay - py, Ex = (long double)bx - px, Ey = (long double)by - py,
                    Fx = (long double)cx - px, Fy = (long double)cy - py;
  const long double Ap = Dx * Dx + Dy * Dy, Bp = Ex * Ex + Ey * Ey, Cp = Fx * Fx + Fy * Fy;
  return (double)(Dx * (Ey * Cp - Bp * Fy) - Dy * (Ex * Cp - Bp * Fx) + Ap * (Ex * Fy - Ey * Fx));
}

struct Builder {
  const double *xy;
  int n;
  std::vector<int32_t> tri;   // 3 vertex ids per triangle
  std::vector<int32_t> half;  // opposite half-edge of each triangle edge, -1 on the hull
  std::vector<int32_t> hull_prev, hull_next, hull_tri, hull_hash;
  int hash_size = 0;
  double cx = 0, cy = 0;
  std::vector<int32_t> stack;

  double px(int i) const { return xy[2 * (size_t)i]; }
  double py(int i) const { return xy[2 * (size_t)i + 1]; }

  int hash_key(double x, double y) const {
    const double dx = x - cx, dy = y - cy;
    const double p = dx / (std::fabs(dx) + std::fabs(dy));  // monotone pseudo-angle in [0, 1), running with the hull order
    const double a = (dy < 0 ? 3.0 - p : 1.0 + p) / 4.0;
    if (!(a == a)) return 0;                                // the point is the centre itself
    int k = (int)std::floor(a * hash_size);
    return k >= hash_size ? k % hash_size : (k < 0 ? 0 : k);
  }
  void link(int a, int b) {
    half[a] = b;
    if (b >= 0) half[b] = a;
  }
  int add_triangle(int i0, int i1, int i2, int a, int b, int c) {
    const int t = (int)tri.size();
    tri.push_back(i0); tri.push_back(i1); tri.push_back(i2);
    half.push_back(-1); half.push_back(-1); half.push_back(-1);
    link(t, a); link(t + 1, b); link(t + 2, c);
    return t;
  }
  // flip edges until the triangles around half-edge `a` are locally Delaunay; returns the hull-facing edge
  int legalize(int a) {
    stack.clear();
    int ar = 0;
    for (;;) {
      const int b = half[a];
      const int a0 = a - a % 3;
      ar = a0 + (a + 2) % 3;
      if (b < 0) {
        if (stack.empty()) break;
        a = stack.back(); stack.pop_back();
        continue;
      }
      const int b0 = b - b % 3;
      const int al = a0 + (a + 1) % 3, bl = b0 + (b + 2) % 3;
      const int p0 = tri[ar], pr = tri[a], pl = tri[al], p1 = tri[bl];
      // triangle (p0, pr, pl) is counter-clockwise; flip when p1 lies strictly inside its circumcircle
      const bool illegal = in_circle(px(p0), py(p0), px(pr), py(pr), px(pl), py(pl), px(p1), py(p1)) > 0;
      if (illegal) {
        tri[a] = p1;
        tri[b] = p0;
        const int hbl = half[bl];
        if (hbl < 0) {  // the flipped edge was on the hull: fix the hull's triangle reference
          int e = hull_start;
          do {
            if (hull_tri[e] == bl) { hull_tri[e] = a; break; }
            e = hull_prev[e];
          } while (e != hull_start);
        }
        link(a, hbl);
        link(b, half[ar]);
        link(ar, bl);
        const int br = b0 + (b + 1) % 3;
        stack.push_back(br);
      } else {
        if (stack.empty()) break;
        a = stack.back(); stack.pop_back();
      }
    }
    return ar;
  }
  int hull_start = 0;
};

}  // namespace

int delaunay_triangulate(const double *xy, int n, std::vector<int32_t> &triangles) {
  triangles.clear();
  if (n < 3) return 0;
  for (int i = 0; i < 2 * n; i++)
    if (!std::isfinite(xy[i])) return -1;
  Builder B;
  B.xy = xy; B.n = n;
  // seed: the point closest to the bounding-box centre, its nearest neighbour, and the third point that
  // gives the smallest circumcircle
  double minx = HUGE_VAL, miny = HUGE_VAL, maxx = -HUGE_VAL, maxy = -HUGE_VAL;
  for (int i = 0; i < n; i++) {
    minx = std::min(minx, B.px(i)); maxx = std::max(maxx, B.px(i));
    miny = std::min(miny, B.py(i)); maxy = std::max(maxy, B.py(i));
  }
  const double bx = 0.5 * (minx + maxx), by = 0.5 * (miny + maxy);
  auto d2 = [&](double ax, double ay, double cx2, double cy2) { return (ax - cx2) * (ax - cx2) + (ay - cy2) * (ay - cy2); };
  int i0 = 0, i1 = -1, i2 = -1;
  double best = HUGE_VAL;
  for (int i = 0; i < n; i++) { const double d = d2(B.px(i), B.py(i), bx, by); if (d < best) { best = d; i0 = i; } }
  best = HUGE_VAL;
  for (int i = 0; i < n; i++) {
    if (i == i0) continue;
    const double d = d2(B.px(i), B.py(i), B.px(i0), B.py(i0));
    if (d < best && d > 0) { best = d; i1 = i; }
  }
  if (i1 < 0) return 0;  // all points coincide
  auto circumradius2 = [&](int a, int b, int c) {
    const double dx = B.px(b) - B.px(a), dy = B.py(b) - B.py(a), ex = B.px(c) - B.px(a), ey = B.py(c) - B.py(a);
    const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = 0.5 / (dx * ey - dy * ex);
    const double x = (ey * bl - dy * cl) * d, y = (dx * cl - ex * bl) * d;
    return x * x + y * y;
  };
  best = HUGE_VAL;
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1) continue;
    const double r = circumradius2(i0, i1, i);
    if (r < best) { best = r; i2 = i; }   // NaN / inf (collinear) never wins
  }
  if (i2 < 0 || !std::isfinite(best)) return 0;  // all points collinear: qhull reports no Delaunay facet either
  if (orient(B.px(i0), B.py(i0), B.px(i1), B.py(i1), B.px(i2), B.py(i2)) < 0) std::swap(i1, i2);
  {
    const double ax = B.px(i0), ay = B.py(i0);
    const double dx = B.px(i1) - ax, dy = B.py(i1) - ay, ex = B.px(i2) - ax, ey = B.py(i2) - ay;
    const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = 0.5 / (dx * ey - dy * ex);
    B.cx = ax + (ey * bl - dy * cl) * d;
    B.cy = ay + (dx * cl - ex * bl) * d;
  }
  std::vector<int32_t> ids(n);
  std::vector<double> dist(n);
  for (int i = 0; i < n; i++) { ids[i] = i; dist[i] = d2(B.px(i), B.py(i), B.cx, B.cy); }
  std::sort(ids.begin(), ids.end(), [&](int a, int b) { return dist[a] < dist[b] || (dist[a] == dist[b] && a < b); });

  B.hash_size = (int)std::ceil(std::sqrt((double)n));
  B.hull_prev.assign(n, 0); B.hull_next.assign(n, 0); B.hull_tri.assign(n, 0); B.hull_hash.assign(B.hash_size, -1);
  B.tri.reserve(6 * (size_t)n); B.half.reserve(6 * (size_t)n);
  B.hull_start = i0;
  B.hull_next[i0] = B.hull_prev[i2] = i1;
  B.hull_next[i1] = B.hull_prev[i0] = i2;
  B.hull_next[i2] = B.hull_prev[i1] = i0;
  B.hull_tri[i0] = 0; B.hull_tri[i1] = 1; B.hull_tri[i2] = 2;
  B.hull_hash[B.hash_key(B.px(i0), B.py(i0))] = i0;
  B.hull_hash[B.hash_key(B.px(i1), B.py(i1))] = i1;
  B.hull_hash[B.hash_key(B.px(i2), B.py(i2))] = i2;
  B.add_triangle(i0, i1, i2, -1, -1, -1);

  double xp = 0, yp = 0;
  bool have_prev = false;
  for (int k = 0; k < n; k++) {
    const int i = ids[k];
    const double x = B.px(i), y = B.py(i);
    if (have_prev && x == xp && y == yp) continue;   // coincident with the previous point (qhull's Qc keeps no vertex for it)
    xp = x; yp = y; have_prev = true;
    if (i == i0 || i == i1 || i == i2) continue;
    // a hull edge visible from the point, starting from the hash bucket of its pseudo-angle
    int start = 0;
    const int key = B.hash_key(x, y);
    for (int j = 0; j < B.hash_size; j++) {
      start = B.hull_hash[(key + j) % B.hash_size];
      if (start != -1 && start != B.hull_next[start]) break;
    }
    start = B.hull_prev[start];
    int e = start, q;
    bool found = true;
    while (q = B.hull_next[e], orient(x, y, B.px(e), B.py(e), B.px(q), B.py(q)) >= 0) {
      e = q;
      if (e == start) { found = false; break; }
    }
    if (!found) continue;  // (numerically) inside the hull and on no visible edge: a duplicate of a hull vertex
    // first triangle from the point
    int t = B.add_triangle(e, i, B.hull_next[e], -1, -1, B.hull_tri[e]);
    B.hull_tri[i] = B.legalize(t + 2);
    B.hull_tri[e] = t;
    // walk forward through the hull, adding triangles while the edges are visible
    int nx = B.hull_next[e];
    while (q = B.hull_next[nx], orient(x, y, B.px(nx), B.py(nx), B.px(q), B.py(q)) < 0) {
      t = B.add_triangle(nx, i, q, B.hull_tri[i], -1, B.hull_tri[nx]);
      B.hull_tri[i] = B.legalize(t + 2);
      B.hull_next[nx] = nx;  // removed from the hull
      nx = q;
    }
    // and backward
    if (e == start) {
      while (q = B.hull_prev[e], orient(x, y, B.px(q), B.py(q), B.px(e), B.py(e)) < 0) {
        t = B.add_triangle(q, i, e, -1, B.hull_tri[e], B.hull_tri[q]);
        B.legalize(t + 2);
        B.hull_tri[q] = t;
        B.hull_next[e] = e;  // removed from the hull
        e = q;
      }
    }
    B.hull_start = B.hull_prev[i] = e;
    B.hull_next[e] = B.hull_prev[nx] = i;
    B.hull_next[i] = nx;
    B.hull_hash[B.hash_key(x, y)] = i;
    B.hull_hash[B.hash_key(B.px(e), B.py(e))] = e;
  }
  triangles.swap(B.tri);
  return (int)(triangles.size() / 3);
}

}  // namespace spr
