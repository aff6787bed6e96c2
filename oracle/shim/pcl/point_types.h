// oracle/shim/pcl/point_types.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for <pcl/point_types.h> so that the reference's
// include/ikd-Tree/ikd_Tree.{h,cpp} can be compiled verbatim (from where it
// lies under /root/reference) without PCL installed. Only the type the tree
// uses is provided: pcl::PointXYZINormal with PCL's documented memory layout
// (48 bytes, 16-byte aligned: {x,y,z,1 | normal_x,normal_y,normal_z,0 |
// intensity,curvature,pad,pad}), zero-initialised like PCL's constructor.
#pragma once
#include <unistd.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZINormal {
    float x, y, z, data3;
    float normal_x, normal_y, normal_z, data_n3;
    float intensity, curvature, pad0, pad1;
    PointXYZINormal()
        : x(0.f), y(0.f), z(0.f), data3(1.f),
          normal_x(0.f), normal_y(0.f), normal_z(0.f), data_n3(0.f),
          intensity(0.f), curvature(0.f), pad0(0.f), pad1(0.f) {}
};
}  // namespace pcl
