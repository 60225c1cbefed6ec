// spr_delaunay.h -- host Delaunay triangulation (spr_delaunay.cpp), the triangle source of SlideGraph.
#pragma once
#include <cstdint>
#include <vector>

namespace spr {
// xy: n x 2.  triangles: 3 vertex ids per triangle, counter-clockwise.  Returns the triangle count
// (0: fewer than 3 distinct, non-collinear points), -1: non-finite coordinate.
int delaunay_triangulate(const double *xy, int n, std::vector<int32_t> &triangles);
}  // namespace spr
