// ref_outline_shim.cpp — C entry points over the REFERENCE'S OWN marching-squares / Douglas-Peucker code, for pinning the oracle.
// TEST INFRASTRUCTURE ONLY.  Makefile.ref compiles /root/reference/source/engine/physics/physics_math.cpp where it lies (with the
// stub SDL headers of oracle/ref_stubs/: the file includes engine/core/sdl_wrapper.h but uses nothing from SDL) together with this
// shim into oracle/_ref/libfse_ref_outline.so.  No reference source is copied; this file only calls the reference's declarations:
//   ME::MarchingSquares::value / FindPerimeter   physics/physics_math.cpp:1870-1965
//   ME::simplify / ME::pDistance                  physics/physics_math.cpp:1766-1843
//   ME::TPPLPoly / ME::TPPLPartition               physics/physics_math.cpp:160-580 (RemoveHoles, Triangulate_EC)
#include <cstdint>
#include <list>
#include <vector>

#include "engine/physics/physics_math.hpp"

#define RAPI extern "C" __attribute__((visibility("default")))

RAPI int ref_ms_value(int x, int y, int w, int h, unsigned char* data) { return ME::MarchingSquares::value(x, y, w, h, data); }

// directions as (x, y) int pairs; returns their count (or -1 when cap is too small); *ix, *iy = Result::initialX / initialY
RAPI int ref_find_perimeter(int x, int y, int w, int h, unsigned char* data, int32_t* dirs, int cap, int32_t* ix, int32_t* iy) {
    ME::MarchingSquares::Result r = ME::MarchingSquares::FindPerimeter(x, y, w, h, data);
    *ix = r.initialX;
    *iy = r.initialY;
    if (r.directions.size() > static_cast<size_t>(cap)) return -1;
    for (size_t i = 0; i < r.directions.size(); i++) {
        dirs[2 * i] = r.directions[i].x;
        dirs[2 * i + 1] = r.directions[i].y;
    }
    const size_t n_dir = r.directions.size();
    return static_cast<int>(n_dir);
}

RAPI int ref_simplify(const float* pts, int n, float tolerance, float* out) {
    std::vector<ME::MEvec2> in_pts;
    for (int i = 0; i < n; i++) in_pts.emplace_back(pts[2 * i], pts[2 * i + 1]);
    std::vector<ME::MEvec2> out_pts = ME::simplify(in_pts, tolerance);
    for (size_t i = 0; i < out_pts.size(); i++) {
        out[2 * i] = out_pts[i].x;
        out[2 * i + 1] = out_pts[i].y;
    }
    const size_t n_out = out_pts.size();
    return static_cast<int>(n_out);
}

RAPI float ref_pdistance(float x, float y, float x1, float y1, float x2, float y2) { return ME::pDistance(x, y, x1, y1, x2, y2); }

// world.cpp:497-557 around the reference's own TPPL: the simplified outlines of one mask (pts / pt_off, tracing order) -> polygons
// (reversed, clockwise = hole) -> RemoveHoles -> Triangulate_EC per polygon -> triangles without the degenerate ones, grouped per
// polygon that kept any.  tris: 6 doubles per triangle.  Returns the number of groups, -1 when a capacity is too small.
RAPI int ref_hitbox_triangles(const float* pts, const int32_t* pt_off, int n_contours, double* tris, int cap_tris, int32_t* group_off, int cap_groups) {
    std::list<ME::TPPLPoly> shapes;
    for (int c = 0; c < n_contours; c++) {
        const int n = pt_off[c + 1] - pt_off[c];
        if (n < 3) continue;
        ME::TPPLPoly poly;
        poly.Init(n);
        for (int i = 0; i < n; i++) poly[n - i - 1] = {pts[2 * (pt_off[c] + i)], pts[2 * (pt_off[c] + i) + 1]};
        if (poly.GetOrientation() == TPPL_CW) poly.SetHole(true);
        if (poly.GetNumPoints() > 2) shapes.push_back(poly);
    }
    std::list<ME::TPPLPoly> result2;
    ME::TPPLPartition part, part2;
    part.RemoveHoles(&shapes, &result2);
    int groups = 0, t = 0;
    for (auto it = result2.begin(); it != result2.end(); it++) {
        std::list<ME::TPPLPoly> result;
        std::list<ME::TPPLPoly> l = {*it};
        part2.Triangulate_EC(&l, &result);
        const int t0 = t;
        for (auto& cur : result) {
            if ((cur[0].x == cur[1].x && cur[1].x == cur[2].x) || (cur[0].y == cur[1].y && cur[1].y == cur[2].y)) continue;
            if (t >= cap_tris) return -1;
            for (int k = 0; k < 3; k++) {
                tris[6 * t + 2 * k] = cur[k].x;
                tris[6 * t + 2 * k + 1] = cur[k].y;
            }
            t++;
        }
        if (t > t0) {
            if (groups >= cap_groups) return -1;
            group_off[groups++] = t0;
        }
    }
    group_off[groups] = t;
    return groups;
}
