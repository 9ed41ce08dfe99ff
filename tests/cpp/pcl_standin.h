// pcl_standin.h — the two PCL types the scan_descriptor interface mentions, with PCL's memory
// layout (pcl::PointXYZI is 32 bytes: x y z pad | intensity pad pad pad), and the reference's
// abstract base restated from its six signatures (descriptor.h:21-36), so that
// include/descriptor_b200.h can be compile-checked and run here without PCL/ROS.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>
namespace pcl {
struct alignas(16) PointXYZI { float x, y, z, _pad; float intensity, _p1, _p2, _p3; };
static_assert(sizeof(PointXYZI) == 32, "PCL layout");
template <typename P> struct PointCloud { std::vector<P> points; };
}
class scan_descriptor
{
public:
	virtual std::vector<float> makeAndSaveDescriptorAndKey(const pcl::PointCloud<pcl::PointXYZI>& scan, const int8_t robot, const int index) = 0;
	virtual void saveDescriptorAndKey(const float* descriptorMat, const int8_t robot, const int index) = 0;
	virtual std::pair<int, float> detectIntraLoopClosureID(const int currentPtr) = 0;
	virtual std::pair<int, float> detectInterLoopClosureID(const int currentPtr) = 0;
	virtual std::pair<int8_t, int> getIndex(const int key) = 0;
	virtual int getSize(const int idIn = -1) = 0;
};
