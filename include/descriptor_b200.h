/*
 * descriptor_b200.h — header-only drop-in for the reference's Scan Context descriptor class.
 *
 *   class scan_context_descriptor_b200 : public scan_descriptor
 *
 * implements the six pure virtuals of class scan_descriptor
 * (/root/reference/include/descriptor.h:21-36) on top of the C-ABI in scl_engine.h, with the
 * constructor signature of scan_context_descriptor (descriptor.h:1307-1316), so that
 * distributedMapping.h:404
 *       scanDescriptor = unique_ptr<scan_descriptor>(new scan_context_descriptor());
 * becomes
 *       scanDescriptor = unique_ptr<scan_descriptor>(new scan_context_descriptor_b200());
 * and nothing else in distributed_mapping changes (call sites :627, :1002, :1072, :1078, :1274,
 * :1280-1284). Link with -lscl_b200.
 *
 * Include AFTER the header that declares `scan_descriptor` and pcl::PointCloud<pcl::PointXYZI>
 * (i.e. after descriptor.h's own includes). Error behaviour mirrors the reference, which has no
 * status channel: "no loop" is first == -1; an engine error is reported on stderr and returns
 * the same "no loop" / empty values. Unlike the reference, getIndex() tolerates out-of-range
 * keys (the reference evaluates getIndex(-1) at distributedMapping.h:1282): it returns {-1,-1}.
 * The class is internally synchronised, so the unlocked queries of loopClosureThread
 * (distributedMapping.h:1078,1280) are safe against concurrent inserts (:625-628, :1001-1003).
 */
#ifndef DESCRIPTOR_B200_H_
#define DESCRIPTOR_B200_H_

#include <cstdint>
#include <cstdio>
#include <utility>
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "scl_engine.h"

class scan_context_descriptor_b200 : public scan_descriptor
{
public:
	scan_context_descriptor_b200(
		int numRing 			= 20,
		int numSector 			= 60,
		int numCandidates 		= 3,
		double distThres 		= 0.14,
		double lidarHeight 		= 1.65,
		double maxRadius 		= 80.0,
		int numExcludeRecent 	= 100,
		int treeMakingPeriod 	= 10,
		double searchRatio 		= 0.1,
		int cudaDevice 			= 0) : engine_(nullptr)
	{
		scl_params p;
		p.num_ring = numRing; p.num_sector = numSector; p.num_candidates = numCandidates;
		p.dist_thres = distThres; p.lidar_height = lidarHeight; p.max_radius = maxRadius;
		p.num_exclude_recent = numExcludeRecent; p.tree_making_period = treeMakingPeriod; p.search_ratio = searchRatio;
		rs_ = numRing * numSector;
		const int rc = scl_create(&p, cudaDevice, &engine_);
		if(rc != SCL_OK)
		{
			std::fprintf(stderr, "[scan_context_descriptor_b200] scl_create failed (%d): a CUDA device is required, there is no CPU fallback\n", rc);
			engine_ = nullptr;
		}
	}

	~scan_context_descriptor_b200()
	{
		if(engine_) scl_destroy(engine_);
	}

	scan_context_descriptor_b200(const scan_context_descriptor_b200&) = delete;
	scan_context_descriptor_b200& operator=(const scan_context_descriptor_b200&) = delete;

	/* descriptor.h:1604-1611 — the returned vector is global_descriptor.msg `values` (row-major R*S) */
	std::vector<float> makeAndSaveDescriptorAndKey(const pcl::PointCloud<pcl::PointXYZI>& scan, const int8_t robot, const int index)
	{
		std::vector<float> vT(rs_, 0.0f);
		const void* pts = scan.points.empty() ? nullptr : static_cast<const void*>(&scan.points[0]);
		if(!check(scl_build_insert(engine_, pts, (int)scan.points.size(), (int)sizeof(pcl::PointXYZI), robot, index, vT.data()), "makeAndSaveDescriptorAndKey"))
			keep_keys_dense(vT, robot, index);
		return vT;
	}

	/* distributedMapping.h:996-1003 as one call — downSizeFilterDes (a pcl::VoxelGrid with a cubic leaf) followed by
	 * makeAndSaveDescriptorAndKey; the filtered cloud stays on the device. *filteredSize = points after the filter. */
	std::vector<float> makeAndSaveDescriptorAndKeyFiltered(const pcl::PointCloud<pcl::PointXYZI>& scan, const float leaf,
		const int8_t robot, const int index, int* filteredSize = nullptr)
	{
		std::vector<float> vT(rs_, 0.0f);
		const void* pts = scan.points.empty() ? nullptr : static_cast<const void*>(&scan.points[0]);
		int m = 0;
		if(!check(scl_build_insert_filtered(engine_, pts, (int)scan.points.size(), (int)sizeof(pcl::PointXYZI), leaf, robot, index, vT.data(), &m),
			"makeAndSaveDescriptorAndKeyFiltered"))
			keep_keys_dense(vT, robot, index);
		if(filteredSize) *filteredSize = m;
		return vT;
	}

	/* descriptor.h:1572-1585 — descriptorMat is msg->values.data(), borrowed for the call */
	void saveDescriptorAndKey(const float* descriptorMat, const int8_t robot, const int index)
	{
		check(scl_insert(engine_, descriptorMat, robot, index), "saveDescriptorAndKey");
	}

	/* descriptor.h:1613-1674 — {loop id or -1, best shift as float} */
	std::pair<int, float> detectIntraLoopClosureID(const int currentPtr)
	{
		int id = -1; float second = 0.0f;
		check(scl_query_intra(engine_, currentPtr, &id, &second), "detectIntraLoopClosureID");
		return std::make_pair(id, second);
	}

	/* descriptor.h:1676-1756 — {loop id or -1, relative yaw in radians} */
	std::pair<int, float> detectInterLoopClosureID(const int currentPtr)
	{
		int id = -1; float second = 0.0f;
		check(scl_query_inter(engine_, currentPtr, &id, &second), "detectInterLoopClosureID");
		return std::make_pair(id, second);
	}

	/* descriptor.h:1758-1761 */
	std::pair<int8_t, int> getIndex(const int key)
	{
		int8_t robot = -1; int index = -1;
		check(scl_get_index(engine_, key, &robot, &index), "getIndex");
		return std::make_pair(robot, index);
	}

	/* descriptor.h:1763-1766 (idIn is ignored there too) */
	int getSize(const int idIn = -1)
	{
		(void)idIn;
		return engine_ ? scl_size(engine_) : 0;
	}

	/* geometric verification for performIntraLoopClosure (distributedMapping.h:1108-1132):
	 * T is the row-major 4x4 of icp.getFinalTransformation() */
	bool icpAlign(const pcl::PointCloud<pcl::PointXYZI>& source, const pcl::PointCloud<pcl::PointXYZI>& target,
		float T[16], float& fitnessScore, double maxCorrespondenceDistance = 100, int maximumIterations = 50,
		double transformationEpsilon = 1e-6, double euclideanFitnessEpsilon = 1e-6)
	{
		scl_icp_params p;
		p.max_corr_dist = maxCorrespondenceDistance; p.max_iterations = maximumIterations;
		p.trans_eps = transformationEpsilon; p.fitness_eps = euclideanFitnessEpsilon;
		int converged = 0, iterations = 0;
		const int rc = scl_icp(engine_, source.points.empty() ? nullptr : &source.points[0], (int)source.points.size(),
			target.points.empty() ? nullptr : &target.points[0], (int)target.points.size(), (int)sizeof(pcl::PointXYZI),
			&p, T, &fitnessScore, &converged, &iterations);
		check(rc, "icpAlign");
		return rc == SCL_OK && converged != 0;
	}

	/* inter-robot verification for geometricVerificationService (distributedMapping.h:1211-1243): nearest-neighbour
	 * correspondences, RANSAC over three-point hypotheses, SVD on the inliers. Returns the `success` of :1238;
	 * T is the row-major 4x4 `transform` of :1228-1230. */
	bool ransacVerify(const pcl::PointCloud<pcl::PointXYZI>& source, const pcl::PointCloud<pcl::PointXYZI>& target,
		float T[16], int ransacMaxIter = 1000, double ransacOutlierTreshold = 0.25, double inlierTreshold = 0.45,
		int* numCorrespondences = nullptr, int* numInliers = nullptr, unsigned seed = 1)
	{
		scl_ransac_params p;
		p.max_iterations = ransacMaxIter; p.inlier_threshold = ransacOutlierTreshold; p.min_inlier_ratio = inlierTreshold; p.seed = seed;
		int nc = 0, ni = 0, ok = 0;
		const int rc = scl_verify_ransac(engine_, source.points.empty() ? nullptr : &source.points[0], (int)source.points.size(),
			target.points.empty() ? nullptr : &target.points[0], (int)target.points.size(), (int)sizeof(pcl::PointXYZI),
			&p, T, &nc, &ni, &ok);
		check(rc, "ransacVerify");
		if(numCorrespondences) *numCorrespondences = nc;
		if(numInliers) *numInliers = ni;
		return rc == SCL_OK && ok != 0;
	}

	/* downSizeFilterDes / downSizeFilterICP (distributedMapping.h:996-998, 1181-1185): pcl::VoxelGrid with a cubic leaf.
	 * `out` receives the centroids (x, y, z, intensity; the padding fields are zeroed). */
	bool voxelGrid(const pcl::PointCloud<pcl::PointXYZI>& cloud, float leaf, pcl::PointCloud<pcl::PointXYZI>& out)
	{
		std::vector<float> buf(cloud.points.size() * 4 + 4);
		int m = 0;
		const int rc = scl_voxel_grid(engine_, cloud.points.empty() ? nullptr : &cloud.points[0], (int)cloud.points.size(),
			(int)sizeof(pcl::PointXYZI), leaf, buf.data(), &m);
		check(rc, "voxelGrid");
		unpack(buf, rc == SCL_OK ? m : 0, out);
		return rc == SCL_OK;
	}

	/* loopFindNearKeyframes (distributedMapping.h:1163-1186): the keyframe clouds moved by their 6-DoF poses
	 * (x, y, z, roll, pitch, yaw: PointPose6D), concatenated and down-sampled (leaf <= 0: not down-sampled). */
	bool assembleSubmap(const std::vector<const pcl::PointCloud<pcl::PointXYZI>*>& clouds, const std::vector<float>& poses6,
		float leaf, pcl::PointCloud<pcl::PointXYZI>& out)
	{
		std::vector<pcl::PointXYZI> all;
		std::vector<int> offsets(1, 0);
		for(size_t c = 0; c < clouds.size(); c++)
		{
			all.insert(all.end(), clouds[c]->points.begin(), clouds[c]->points.end());
			offsets.push_back((int)all.size());
		}
		std::vector<float> buf(all.size() * 4 + 4);
		int m = 0;
		const int rc = scl_assemble_submap(engine_, all.empty() ? nullptr : &all[0], offsets.data(), (int)clouds.size(),
			(int)sizeof(pcl::PointXYZI), poses6.data(), leaf, buf.data(), &m);
		check(rc, "assembleSubmap");
		unpack(buf, rc == SCL_OK ? m : 0, out);
		return rc == SCL_OK;
	}

	scl_engine* engine() { return engine_; }

private:
	static void unpack(const std::vector<float>& buf, int m, pcl::PointCloud<pcl::PointXYZI>& out)
	{
		out.points.resize(m);
		for(int i = 0; i < m; i++)
		{
			pcl::PointXYZI p = pcl::PointXYZI();
			p.x = buf[4 * i]; p.y = buf[4 * i + 1]; p.z = buf[4 * i + 2]; p.intensity = buf[4 * i + 3];
			out.points[i] = p;
		}
	}

	bool check(int rc, const char* what)
	{
		if(rc != SCL_OK)
			std::fprintf(stderr, "[scan_context_descriptor_b200] %s failed (%d): %s\n", what, rc, engine_ ? scl_last_error(engine_) : "no engine");
		return rc == SCL_OK;
	}

	/* The caller takes key = cloudKeyPoses6D->size() - 1 for granted (distributedMapping.h:1002,1072): a build that failed
	 * must still occupy its key, or every later key drifts by one. An empty (all-zero) descriptor takes the slot: it matches
	 * nothing (every SC distance against it is NaN, descriptor.h:1523,1534). If even that fails the process cannot continue. */
	void keep_keys_dense(std::vector<float>& vT, const int8_t robot, const int index)
	{
		std::fill(vT.begin(), vT.end(), 0.0f);
		if(!engine_ || scl_insert(engine_, vT.data(), robot, index) != SCL_OK)
		{
			std::fprintf(stderr, "[scan_context_descriptor_b200] cannot keep the key space dense after a failed build: aborting\n");
			std::abort();
		}
	}

	scl_engine* engine_;
	int rs_;
};

/*
 * The same six virtuals with the keyframe database sharded by keyframe index over several GPUs of the box
 * (scl_create_sharded, include/scl_engine.h): still ONE object behind distributed_mapping's
 * std::unique_ptr<scan_descriptor> (distributedMapping.h:333,402-405). devices = CUDA device ordinals; keys, ids and
 * results mean exactly what they mean in scan_context_descriptor_b200.
 */
class scan_context_descriptor_b200_sharded : public scan_descriptor
{
public:
	scan_context_descriptor_b200_sharded(const std::vector<int>& devices, int numRing = 20, int numSector = 60, int numCandidates = 3,
		double distThres = 0.14, double lidarHeight = 1.65, double maxRadius = 80.0, int numExcludeRecent = 100,
		int treeMakingPeriod = 10, double searchRatio = 0.1) : sharded_(nullptr), rs_(numRing * numSector)
	{
		scl_params p;
		scl_default_params(&p);
		p.num_ring = numRing; p.num_sector = numSector; p.num_candidates = numCandidates; p.dist_thres = distThres;
		p.lidar_height = lidarHeight; p.max_radius = maxRadius; p.num_exclude_recent = numExcludeRecent;
		p.tree_making_period = treeMakingPeriod; p.search_ratio = searchRatio;
		const int rc = scl_create_sharded(&p, (int)devices.size(), devices.data(), 1024, numCandidates < 10 ? 10 : numCandidates, &sharded_);
		if(rc != SCL_OK)
		{
			std::fprintf(stderr, "[scan_context_descriptor_b200_sharded] scl_create_sharded failed (%d): CUDA devices with peer access are required, there is no CPU fallback\n", rc);
			std::abort();
		}
	}

	~scan_context_descriptor_b200_sharded() { if(sharded_) scl_sharded_destroy(sharded_); }
	scan_context_descriptor_b200_sharded(const scan_context_descriptor_b200_sharded&) = delete;
	scan_context_descriptor_b200_sharded& operator=(const scan_context_descriptor_b200_sharded&) = delete;

	std::vector<float> makeAndSaveDescriptorAndKey(const pcl::PointCloud<pcl::PointXYZI>& scan, const int8_t robot, const int index)
	{
		std::vector<float> vT(rs_, 0.0f);
		const void* pts = scan.points.empty() ? nullptr : static_cast<const void*>(&scan.points[0]);
		if(!check(scl_sharded_build_insert(sharded_, pts, (int)scan.points.size(), (int)sizeof(pcl::PointXYZI), robot, index, vT.data()), "makeAndSaveDescriptorAndKey"))
		{
			/* keep the key space dense (see scan_context_descriptor_b200::keep_keys_dense) */
			std::fill(vT.begin(), vT.end(), 0.0f);
			if(scl_sharded_insert_batch(sharded_, vT.data(), 1, &robot, &index) != SCL_OK) std::abort();
		}
		return vT;
	}

	void saveDescriptorAndKey(const float* descriptorMat, const int8_t robot, const int index)
	{
		check(scl_sharded_insert_batch(sharded_, descriptorMat, 1, &robot, &index), "saveDescriptorAndKey");
	}

	std::pair<int, float> detectIntraLoopClosureID(const int currentPtr)
	{
		int id = -1; float second = 0.0f;
		check(scl_sharded_query_intra(sharded_, currentPtr, &id, &second), "detectIntraLoopClosureID");
		return std::make_pair(id, second);
	}

	std::pair<int, float> detectInterLoopClosureID(const int currentPtr)
	{
		int id = -1; float second = 0.0f;
		check(scl_sharded_query_inter(sharded_, currentPtr, &id, &second), "detectInterLoopClosureID");
		return std::make_pair(id, second);
	}

	std::pair<int8_t, int> getIndex(const int key)
	{
		int8_t robot = -1; int index = -1;
		check(scl_sharded_get_index(sharded_, key, &robot, &index), "getIndex");
		return std::make_pair(robot, index);
	}

	int getSize(const int idIn = -1) { (void)idIn; return scl_sharded_size(sharded_); }

private:
	bool check(int rc, const char* what)
	{
		if(rc != SCL_OK)
			std::fprintf(stderr, "[scan_context_descriptor_b200_sharded] %s failed (%d): %s\n", what, rc, scl_sharded_last_error(sharded_));
		return rc == SCL_OK;
	}

	scl_sharded* sharded_;
	int rs_;
};

#endif
