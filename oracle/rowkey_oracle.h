/*
 * rowkey_oracle.h — C entry points of the CPU ORACLE for the row-key candidate stage of the reference's
 * class lidar_iris_descriptor (/root/reference/include/descriptor.h:1047-1059 save, :1087-1148
 * detectIntraLoopClosureID, :1150-1250 detectInterLoopClosureID, :1252-1267 getIndex / getSize).
 *
 * TEST INFRASTRUCTURE ONLY (see sc_oracle.h): two libraries export these symbols,
 *   oracle/liboracle.so            our restatement, rowkey_oracle.cpp
 *   oracle/_ref/libiris_ref.so     the reference's OWN text of those functions and of the class's member list, cut from
 *                                  descriptor.h at build time (oracle/Makefile) into a shell class (iris_ref_driver.cpp)
 * The LiDAR-Iris image features and compare() are OpenCV code (out of scope, absent here): in both libraries a stored
 * entry carries one float `feature` instead, and compare(a, b) is |fa - fb| with bias (7 * tag_a + 13 * tag_b) % 360,
 * tag = the entry's global key — enough to exercise the candidate scan, the strict-< minimum and the threshold.
 * libnabo is absent: both use the linear-scan stand-in rules (parity unpinned for exact ties).
 */
#ifndef ROWKEY_ORACLE_H_
#define ROWKEY_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sco_iris sco_iris;
sco_iris* sco_iris_create(int rows, int num_exclude_recent, int num_candidates, double dist_thres, int robot_num, int this_id);
void sco_iris_destroy(sco_iris* h);
void sco_iris_save(sco_iris* h, const float* row_key, int robot, int index, float feature);
/* result of detect*LoopClosureID; n_cand / cand / cand_d2 (num_candidates entries): the kNN list of the call, 0 / untouched
 * when the function returned before the search */
void sco_iris_detect_intra(sco_iris* h, int cur_ptr, int* id, float* bias, int* n_cand, int32_t* cand, float* cand_d2);
void sco_iris_detect_inter(sco_iris* h, int cur_ptr, int* id, float* bias, int* n_cand, int32_t* cand, float* cand_d2);
void sco_iris_get_index(sco_iris* h, int key, int* robot, int* index);
int sco_iris_size(sco_iris* h, int id_in);
/* timing aid: the kNN alone of Q query keys against robot's first n keys, `threads` host threads */
void sco_iris_knn_batch(sco_iris* h, const float* q_keys, int Q, int robot, int n, int K, int threads, int32_t* idx, float* d2);
#ifdef __cplusplus
}
#endif
#endif
