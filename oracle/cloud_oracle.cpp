// cloud_oracle.cpp — TEST INFRASTRUCTURE, not product: CPU restatement of the cloud preparation around the descriptor
// and ICP. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use it.
//
//   sco_voxel_grid_pcl    pcl::VoxelGrid<PointXYZI>::applyFilter as the reference drives it
//                         (/root/reference/include/distributedMapping.h:996-998 downSizeFilterDes,
//                          :1181-1185 and :1200-1201 downSizeFilterICP; leaf sizes set at :501-503)
//   sco_assemble_submap   loopFindNearKeyframes (:1163-1186) = transformPointCloud (:234-253) per keyframe,
//                         concatenation, VoxelGrid
//
// PARITY UNPINNED for the VoxelGrid half: PCL (zhongshp5/pcl_catkin @ afe789a, dependencies.rosinstall:33-36) is not
// vendored in /root/reference and the reference holds no test for it. The code below restates PCL's published
// algorithm (filters/impl/voxel_grid.hpp applyFilter + common/centroid.h CentroidPoint, float arithmetic):
//   min/max of the finite points; inverse leaf = 1.0f / leaf; refuse (return the input) if the leaf grid would
//   overflow an int; min_b = int(floor(min * inv)); ijk = int(floor(p * inv) - float(min_b));
//   idx = ijk.x + ijk.y * div.x + ijk.z * div.x * div.y; sort by idx; per leaf float sums / float(n); output in idx order.
// PCL sorts with std::sort, whose order among equal keys is unspecified; the oracle DEFINES it as the input order.
// The transformPointCloud half follows the reference's own expression (:246-248) and pcl::getTransformation's
// published formula (common/impl/eigen.hpp), evaluated in float with this libm's sinf/cosf.
#include "sc_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

extern "C" {

int sco_voxel_grid_pcl(const float* pts, int n, int stride, float leaf, float* out /* 4 floats per point, room for n */)
{
    if (n <= 0) return 0;
    bool any = false;
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) {
        const float* p = pts + (size_t)i * stride;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        if (!any) { for (int a = 0; a < 3; a++) mn[a] = mx[a] = p[a]; any = true; }
        for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
    }
    if (!any) return 0;
    const float inv = 1.0f / leaf;
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)INT32_MAX) {                 /* "Leaf size is too small": output = input */
        for (int i = 0; i < n; i++) memcpy(out + (size_t)i * 4, pts + (size_t)i * stride, 16);
        return n;
    }
    int min_b[3], max_b[3];
    for (int a = 0; a < 3; a++) { min_b[a] = (int)std::floor(mn[a] * inv); max_b[a] = (int)std::floor(mx[a] * inv); }
    const int div0 = max_b[0] - min_b[0] + 1, div1 = max_b[1] - min_b[1] + 1;
    std::vector<std::pair<uint32_t, int>> keyed;
    keyed.reserve(n);
    for (int i = 0; i < n; i++) {
        const float* p = pts + (size_t)i * stride;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        const int i0 = (int)(std::floor(p[0] * inv) - (float)min_b[0]);
        const int i1 = (int)(std::floor(p[1] * inv) - (float)min_b[1]);
        const int i2 = (int)(std::floor(p[2] * inv) - (float)min_b[2]);
        keyed.emplace_back((uint32_t)(i0 + i1 * div0 + i2 * div0 * div1), i);
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) { return a.first < b.first; });
    int m = 0;
    for (size_t i = 0; i < keyed.size();) {
        size_t j = i;
        float sx = 0.0f, sy = 0.0f, sz = 0.0f, si = 0.0f;
        for (; j < keyed.size() && keyed[j].first == keyed[i].first; j++) {
            const float* p = pts + (size_t)keyed[j].second * stride;
            sx += p[0]; sy += p[1]; sz += p[2]; si += p[3];
        }
        const float c = (float)(j - i);
        out[(size_t)m * 4 + 0] = sx / c; out[(size_t)m * 4 + 1] = sy / c; out[(size_t)m * 4 + 2] = sz / c; out[(size_t)m * 4 + 3] = si / c;
        m++; i = j;
    }
    return m;
}

int sco_assemble_submap(const float* pts, const int* offsets, int n_clouds, int stride, const float* poses6, float leaf, float* out)
{
    const int total = n_clouds > 0 ? offsets[n_clouds] : 0;
    if (total <= 0) return 0;
    std::vector<float> world((size_t)total * 4);
    for (int c = 0; c < n_clouds; c++) {
        const float* p = poses6 + (size_t)c * 6;
        /* pcl::getTransformation(x, y, z, roll, pitch, yaw), Scalar = float */
        const float A = std::cos(p[5]), B = std::sin(p[5]), C = std::cos(p[4]), D = std::sin(p[4]), E = std::cos(p[3]), F = std::sin(p[3]);
        const float DE = D * E, DF = D * F;
        const float t[12] = {A * C, A * DF - B * E, B * F + A * DE, p[0],
                             B * C, A * E + B * DF, B * DE - A * F, p[1],
                             -D, C * F, C * E, p[2]};
        for (int i = offsets[c]; i < offsets[c + 1]; i++) {
            const float* q = pts + (size_t)i * stride;
            float* o = &world[(size_t)i * 4];
            o[0] = t[0] * q[0] + t[1] * q[1] + t[2] * q[2] + t[3];       /* distributedMapping.h:246-248, left to right */
            o[1] = t[4] * q[0] + t[5] * q[1] + t[6] * q[2] + t[7];
            o[2] = t[8] * q[0] + t[9] * q[1] + t[10] * q[2] + t[11];
            o[3] = q[3];
        }
    }
    if (!(leaf > 0.0f)) { memcpy(out, world.data(), world.size() * 4); return total; }
    return sco_voxel_grid_pcl(world.data(), total, 4, leaf, out);
}

} // extern "C"
