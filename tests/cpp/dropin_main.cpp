// dropin_main.cpp — uses scan_context_descriptor_b200 exactly the way distributed_mapping uses
// scan_context_descriptor: through unique_ptr<scan_descriptor> (distributedMapping.h:333,404),
// build+insert from a cloud (:1002), insert from a wire vector (:627), intra/inter queries
// (:1078,1280), getIndex/getSize (:1072,1281-1284). Reads a small binary scenario written by
// tests/test_dropin_cpp.py and prints one line per call for the test to compare with the oracle. With a second argument n
// the object is scan_context_descriptor_b200_sharded over n GPUs (scl_create_sharded).
#include "pcl_standin.h"
#include "../../include/descriptor_b200.h"
#include <cstdio>
#include <cstdlib>
#include <memory>

int main(int argc, char** argv)
{
	if(argc < 2) return 2;
	FILE* f = std::fopen(argv[1], "rb");
	if(!f) return 2;
	int n_clouds = 0, n_wires = 0, rs = 0;
	if(std::fread(&n_clouds, 4, 1, f) != 1 || std::fread(&n_wires, 4, 1, f) != 1 || std::fread(&rs, 4, 1, f) != 1) return 2;
	/* argv[2] = number of GPUs: the sharded adapter over devices 0 .. n-1 (still one object behind the same pointer) */
	std::unique_ptr<scan_descriptor> scanDescriptor;
	if(argc >= 3)
	{
		std::vector<int> devices;
		for(int d = 0; d < std::atoi(argv[2]); d++) devices.push_back(d);
		scanDescriptor.reset(new scan_context_descriptor_b200_sharded(devices, 20, 60, 10, 0.14, 1.65, 80.0, 30));
	}
	else scanDescriptor.reset(new scan_context_descriptor_b200(20, 60, 10, 0.14, 1.65, 80.0, 30));
	for(int i = 0; i < n_clouds; i++)
	{
		int np = 0;
		if(std::fread(&np, 4, 1, f) != 1) return 2;
		pcl::PointCloud<pcl::PointXYZI> cloud;
		cloud.points.resize(np);
		if(np && std::fread(cloud.points.data(), sizeof(pcl::PointXYZI), np, f) != (size_t)np) return 2;
		std::vector<float> v = scanDescriptor->makeAndSaveDescriptorAndKey(cloud, 0, i);
		double sum = 0; for(float x : v) sum += x;
		std::printf("build %d %zu %.9g\n", i, v.size(), sum);
	}
	std::vector<float> wire(rs);
	for(int i = 0; i < n_wires; i++)
	{
		if(std::fread(wire.data(), 4, rs, f) != (size_t)rs) return 2;
		scanDescriptor->saveDescriptorAndKey(wire.data(), 1, i);
	}
	std::fclose(f);
	const int n = scanDescriptor->getSize();
	std::printf("size %d\n", n);
	for(int cur = 0; cur < n; cur++)
	{
		std::pair<int, float> a = scanDescriptor->detectIntraLoopClosureID(cur);
		std::pair<int, float> b = scanDescriptor->detectInterLoopClosureID(cur);
		std::printf("query %d %d %.9g %d %.9g\n", cur, a.first, a.second, b.first, b.second);
	}
	std::pair<int8_t, int> ix = scanDescriptor->getIndex(n_clouds);
	std::pair<int8_t, int> bad = scanDescriptor->getIndex(-1);
	std::printf("index %d %d %d %d\n", (int)ix.first, ix.second, (int)bad.first, bad.second);
	return 0;
}
