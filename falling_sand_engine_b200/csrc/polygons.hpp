// polygons.hpp — the host-side half of world::updateRigidBodyHitbox / updateChunkMesh (world.cpp:490-563, 884-940): the outlines the
// device traced (fse_mask_outline) become polygons, holes are bridged into their outer polygons and every outer polygon is cut into
// triangles for b2PolygonShape.  The reference does this with its vendored polypartition (TPPLPartition::RemoveHoles and
// Triangulate_EC, physics/physics_math.cpp:297-580); this is an independent implementation of the same two published algorithms —
// bridge every hole from its right-most vertex to the visible outer vertex with the best bearing; clip the ear with the widest
// opening first — arranged so that it makes the same choices in the same order and evaluates the same double-precision expressions,
// because Box2D fixtures (and the fracture behaviour they give) depend on which triangles come out.  tests/test_polygons.py holds it
// against the reference's own compiled code (oracle/_ref).  Pure host code: no CUDA, no allocation beyond std containers.
#pragma once
#include <cmath>
#include <cstddef>
#include <list>
#include <vector>

namespace fse_host {

struct Vec2d {
    double x = 0, y = 0;
};
struct Polygon {
    std::vector<Vec2d> pts;
    bool hole = false;
};
struct Triangle {
    Vec2d p[3];
};

namespace polydetail {

// twice the signed area, accumulated edge by edge in index order (the sign decides CCW / CW, physics_math.cpp:221-233)
inline int orientation(const std::vector<Vec2d>& v) {
    double area = 0;
    const size_t n = v.size();
    for (size_t i = 0; i < n; i++) {
        const size_t j = i + 1 == n ? 0 : i + 1;
        area += v[i].x * v[j].y - v[i].y * v[j].x;
    }
    return area > 0 ? 1 : (area < 0 ? -1 : 0);
}

// left turn a -> b -> c (strictly), physics_math.cpp:414-421
inline bool left_turn(const Vec2d& a, const Vec2d& b, const Vec2d& c) { return (c.y - a.y) * (b.x - a.x) - (c.x - a.x) * (b.y - a.y) > 0; }

inline Vec2d unit(const Vec2d& d) {
    const double n = std::sqrt(d.x * d.x + d.y * d.y);
    Vec2d r;
    if (n != 0) {
        r.x = d.x / n;
        r.y = d.y / n;
    }
    return r;
}

// is q inside the wedge that the polygon spans at vertex b (neighbours a and c)?  physics_math.cpp:439-453
inline bool in_wedge(const Vec2d& a, const Vec2d& b, const Vec2d& c, const Vec2d& q) {
    if (left_turn(a, b, c)) return left_turn(a, b, q) && left_turn(b, c, q);
    return left_turn(a, b, q) || left_turn(b, c, q);
}

// do the segments (a0, a1) and (b0, b1) cross?  Shared end points do not count.  physics_math.cpp:263-294
inline bool segments_cross(const Vec2d& a0, const Vec2d& a1, const Vec2d& b0, const Vec2d& b1) {
    auto same = [](const Vec2d& u, const Vec2d& v) { return u.x == v.x && u.y == v.y; };
    if (same(a0, b0) || same(a0, b1) || same(a1, b0) || same(a1, b1)) return false;
    const Vec2d na{a1.y - a0.y, a0.x - a1.x}, nb{b1.y - b0.y, b0.x - b1.x};  // normals of the two segments
    const double b0_side = (b0.x - a0.x) * na.x + (b0.y - a0.y) * na.y, b1_side = (b1.x - a0.x) * na.x + (b1.y - a0.y) * na.y;
    const double a0_side = (a0.x - b0.x) * nb.x + (a0.y - b0.y) * nb.y, a1_side = (a1.x - b0.x) * nb.x + (a1.y - b0.y) * nb.y;
    if (a0_side * a1_side > 0) return false;
    if (b0_side * b1_side > 0) return false;
    return true;
}

}  // namespace polydetail

inline int orientation(const Polygon& p) { return polydetail::orientation(p.pts); }

// Every hole is merged into an outer polygon through a two-way bridge, right-most hole vertex first, until none is left
// (RemoveHoles).  Returns false, leaving `out` as it was, when some hole sees no outer vertex to its right.
inline bool remove_holes(const std::list<Polygon>& in, std::list<Polygon>& out) {
    using namespace polydetail;
    bool any = false;
    for (const Polygon& p : in) any = any || p.hole;
    if (!any) {
        for (const Polygon& p : in) out.push_back(p);
        return true;
    }
    std::list<Polygon> work(in);
    for (;;) {
        // the hole vertex with the largest x over all holes; ties keep the earlier one
        auto hole_it = work.end();
        size_t hole_v = 0;
        for (auto it = work.begin(); it != work.end(); ++it) {
            if (!it->hole) continue;
            if (hole_it == work.end()) {
                hole_it = it;
                hole_v = 0;
            }
            for (size_t i = 0; i < it->pts.size(); i++)
                if (it->pts[i].x > hole_it->pts[hole_v].x) {
                    hole_it = it;
                    hole_v = i;
                }
        }
        if (hole_it == work.end()) break;
        const Vec2d hp = hole_it->pts[hole_v];

        // the outer vertex to bridge to: right of the hole vertex, with the hole vertex inside its wedge, visible (the bridge crosses
        // no outer edge), and of all those the one whose direction from the hole vertex has the largest x component
        auto outer_it = work.end();
        size_t outer_v = 0;
        Vec2d best{};
        bool found = false;
        for (auto it = work.begin(); it != work.end(); ++it) {
            if (it->hole) continue;
            const size_t n = it->pts.size();
            for (size_t i = 0; i < n; i++) {
                const Vec2d& cand = it->pts[i];
                if (cand.x <= hp.x) continue;
                if (!in_wedge(it->pts[(i + n - 1) % n], cand, it->pts[(i + 1) % n], hp)) continue;
                if (found) {
                    const Vec2d dc = unit(Vec2d{cand.x - hp.x, cand.y - hp.y}), db = unit(Vec2d{best.x - hp.x, best.y - hp.y});
                    if (db.x > dc.x) continue;
                }
                bool visible = true;
                for (auto e = work.begin(); e != work.end() && visible; ++e) {
                    if (e->hole) continue;
                    const size_t m = e->pts.size();
                    for (size_t k = 0; k < m; k++)
                        if (segments_cross(hp, cand, e->pts[k], e->pts[(k + 1) % m])) {
                            visible = false;
                            break;
                        }
                }
                if (visible) {
                    found = true;
                    best = cand;
                    outer_it = it;
                    outer_v = i;
                }
            }
        }
        if (!found) return false;

        // outer[0 .. v], the hole once around starting and ending at its bridge vertex, outer[v .. end)
        Polygon merged;
        const size_t hn = hole_it->pts.size();
        merged.pts.reserve(hn + outer_it->pts.size() + 2);
        for (size_t i = 0; i <= outer_v; i++) merged.pts.push_back(outer_it->pts[i]);
        for (size_t i = 0; i <= hn; i++) merged.pts.push_back(hole_it->pts[(i + hole_v) % hn]);
        for (size_t i = outer_v; i < outer_it->pts.size(); i++) merged.pts.push_back(outer_it->pts[i]);
        work.erase(hole_it);
        work.erase(outer_it);
        work.push_back(std::move(merged));
    }
    for (Polygon& p : work) out.push_back(std::move(p));
    return true;
}

// Ear clipping (Triangulate_EC): of all current ears the one with the largest cosine between its two edges — the widest opening —
// goes first; ties keep the lowest index.  Triangles are appended to `out`; false when the polygon is degenerate or no ear is left
// (what was appended until then stays, as with the reference, whose caller ignores the return value).
inline bool triangulate_ec(const Polygon& poly, std::vector<Triangle>& out) {
    using namespace polydetail;
    const long n = (long)poly.pts.size();
    if (n < 3) return false;
    if (n == 3) {
        out.push_back(Triangle{{poly.pts[0], poly.pts[1], poly.pts[2]}});
        return true;
    }
    struct Node {
        long prev, next;
        bool active, convex, ear;
        double cosine;
    };
    std::vector<Node> nd((size_t)n);
    const std::vector<Vec2d>& P = poly.pts;
    auto refresh = [&](long i) {
        Node& v = nd[(size_t)i];
        const Vec2d &a = P[(size_t)v.prev], &b = P[(size_t)i], &c = P[(size_t)v.next];
        v.convex = left_turn(a, b, c);
        const Vec2d ua = unit(Vec2d{a.x - b.x, a.y - b.y}), uc = unit(Vec2d{c.x - b.x, c.y - b.y});
        v.cosine = ua.x * uc.x + ua.y * uc.y;
        v.ear = v.convex;
        if (!v.convex) return;
        for (long k = 0; k < n; k++) {  // every vertex of the polygon, clipped or not, that is not one of the three corners
            const Vec2d& q = P[(size_t)k];
            if ((q.x == b.x && q.y == b.y) || (q.x == a.x && q.y == a.y) || (q.x == c.x && q.y == c.y)) continue;
            if (!left_turn(a, q, b) && !left_turn(b, q, c) && !left_turn(c, q, a)) {  // inside or on the border of the ear
                v.ear = false;
                return;
            }
        }
    };
    for (long i = 0; i < n; i++) nd[(size_t)i] = Node{i == 0 ? n - 1 : i - 1, i == n - 1 ? 0 : i + 1, true, false, false, 0.0};
    for (long i = 0; i < n; i++) refresh(i);
    for (long step = 0; step < n - 3; step++) {
        long ear = -1;
        for (long j = 0; j < n; j++) {
            if (!nd[(size_t)j].active || !nd[(size_t)j].ear) continue;
            if (ear < 0 || nd[(size_t)j].cosine > nd[(size_t)ear].cosine) ear = j;
        }
        if (ear < 0) return false;
        Node& e = nd[(size_t)ear];
        out.push_back(Triangle{{P[(size_t)e.prev], P[(size_t)ear], P[(size_t)e.next]}});
        e.active = false;
        nd[(size_t)e.prev].next = e.next;
        nd[(size_t)e.next].prev = e.prev;
        if (step == n - 4) break;
        refresh(e.prev);
        refresh(e.next);
    }
    for (long i = 0; i < n; i++)
        if (nd[(size_t)i].active) {
            out.push_back(Triangle{{P[(size_t)nd[(size_t)i].prev], P[(size_t)i], P[(size_t)nd[(size_t)i].next]}});
            break;
        }
    return true;
}

// world.cpp:497-557 for the simplified outlines of one mask (each a closed polyline of >= 3 points in tracing order): the polygon is
// the outline reversed; a clockwise one is a hole; holes are bridged; every resulting polygon is ear-clipped; triangles whose three
// x or three y coincide are dropped (558); polygons left without a triangle give no group.  One group = the b2PolygonShapes of one
// new body.
inline std::vector<std::vector<Triangle>> hitbox_triangles(const std::vector<std::vector<Vec2d>>& outlines) {
    std::list<Polygon> shapes;
    for (const auto& o : outlines) {
        if (o.size() < 3) continue;
        Polygon p;
        p.pts.assign(o.rbegin(), o.rend());
        p.hole = orientation(p) < 0;
        shapes.push_back(std::move(p));
    }
    std::list<Polygon> solid;
    remove_holes(shapes, solid);
    std::vector<std::vector<Triangle>> groups;
    for (const Polygon& p : solid) {
        std::vector<Triangle> all, good;
        triangulate_ec(p, all);
        for (const Triangle& t : all) {
            if ((t.p[0].x == t.p[1].x && t.p[1].x == t.p[2].x) || (t.p[0].y == t.p[1].y && t.p[1].y == t.p[2].y)) continue;
            good.push_back(t);
        }
        if (!good.empty()) groups.push_back(std::move(good));
    }
    return groups;
}

}  // namespace fse_host
