// fse_polygons.cu — C entry point over polygons.hpp (host code only; no kernel is launched from here): the outlines of one mask ->
// triangle groups, the host half of updateRigidBodyHitbox / updateChunkMesh (world.cpp:497-563) behind fse_mask_outline.
#include <vector>

#include "fse_internal.hpp"
#include "polygons.hpp"

using namespace fse;

extern "C" FSE_API int fse_hitbox_triangles(const float* pts, const int32_t* pt_off, int32_t n_contours, double* tris, int32_t cap_tris,
                                            int32_t* group_off, int32_t cap_groups, int32_t* n_groups) {
    if (!pts || !pt_off || n_contours < 0 || !tris || !group_off || !n_groups || cap_groups < 0 || cap_tris < 0)
        return fail(FSE_EINVAL, "fse_hitbox_triangles: bad argument");
    if (n_contours > 0 && pt_off[0] < 0) return fail(FSE_EINVAL, "fse_hitbox_triangles: negative point offset");
    std::vector<std::vector<fse_host::Vec2d>> outlines((size_t)n_contours);
    for (int c = 0; c < n_contours; c++) {
        if (pt_off[c + 1] < pt_off[c]) return fail(FSE_EINVAL, "fse_hitbox_triangles: point offsets must not decrease");
        for (int q = pt_off[c]; q < pt_off[c + 1]; q++) outlines[(size_t)c].push_back(fse_host::Vec2d{(double)pts[2 * q], (double)pts[2 * q + 1]});
    }
    const std::vector<std::vector<fse_host::Triangle>> groups = fse_host::hitbox_triangles(outlines);
    size_t total = 0;
    for (const auto& g : groups) total += g.size();
    *n_groups = (int32_t)groups.size();
    if ((int64_t)groups.size() > cap_groups || (int64_t)total > cap_tris)
        return fail(FSE_ENOMEM, "fse_hitbox_triangles: %zu groups / %zu triangles exceed the caller's capacity (%d / %d)", groups.size(), total, cap_groups,
                    cap_tris);
    size_t t = 0;
    for (size_t g = 0; g < groups.size(); g++) {
        group_off[g] = (int32_t)t;
        for (const fse_host::Triangle& tr : groups[g]) {
            for (int k = 0; k < 3; k++) {
                tris[6 * t + 2 * k] = tr.p[k].x;
                tris[6 * t + 2 * k + 1] = tr.p[k].y;
            }
            t++;
        }
    }
    group_off[groups.size()] = (int32_t)t;
    return FSE_OK;
}
