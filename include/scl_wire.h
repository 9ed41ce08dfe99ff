/* scl_wire.h — wire codec of the loop-closure path: the three messages the path exchanges, as plain C
 * structs, and the pose arithmetic that fills them. Host-side code (no GPU work), exported by libscl_b200.so.
 *
 *   global_descriptor.msg  /root/reference/msg/global_descriptor.msg:1-8, filled at include/distributedMapping.h:1004-1024
 *   loop_info.msg          /root/reference/msg/loop_info.msg:1-9, filled at distributedMapping.h:1129-1158
 *   geometric_verification.srv  /root/reference/srv/geometric_verification.srv:1-8, response filled at :1244-1259
 *
 * The reference computes these with PCL (getTransformation, getTranslationAndEulerAngles: float), GTSAM (Pose3, Rot3:
 * double) and tf (createQuaternionMsgFromRollPitchYaw: double). None of the three is vendored in /root/reference, so the
 * functions below restate their published formulas; parity is unpinned (tests/test_wire.py checks them against an
 * independent numpy restatement and algebraic identities). */
#ifndef SCL_WIRE_H
#define SCL_WIRE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double tx, ty, tz; double qx, qy, qz, qw; } scl_transform;        /* geometry_msgs/Transform */

typedef struct {                        /* global_descriptor.msg (Header left to the caller) */
    int32_t index;                      /* keyframe index, :1019 */
    scl_transform pre_pose, cur_pose;   /* :1008-1017 */
    const float* values;                /* R*S floats, row-major: what scl_build_insert returns and scl_insert takes */
    int32_t n_values;
} scl_global_descriptor;

typedef struct {                        /* loop_info.msg */
    int32_t robot0, robot1, index0, index1;   /* :1145-1148 */
    float noise;                        /* ICP fitness, :1149 */
    scl_transform bet_pose;             /* :1150-1156 */
} scl_loop_info;

/* (x, y, z, roll, pitch, yaw) -> Transform: translation copied, rotation = tf::createQuaternionMsgFromRollPitchYaw (:1013-1017) */
void scl_wire_pose6_to_transform(const float pose6[6], scl_transform* out);

/* The pose-between of a verified loop (:1129-1141 intra, :1249-1256 inter):
 *   tfWrong   = pcl::getTransformation(pose_cur6)                       (float, :1136 / :1249)
 *   tfCorrect = T_align * tfWrong                                        (float 4x4 product, :1137 / :1250)
 *   (x, y, z, roll, pitch, yaw) = pcl::getTranslationAndEulerAngles(tfCorrect)      (:1138 / :1251)
 *   poseFrom  = Pose3(Rot3::RzRyRx(roll, pitch, yaw), Point3(x, y, z)); poseTo = the same from pose_pre6 (:1139-1140)
 *   bet       = poseFrom.between(poseTo)                                 (:1141 / :1254)
 * T_align: row-major 4x4 (scl_icp / scl_verify_ransac T_out). quat_from_rpy = 0: rotation().toQuaternion() as loop_info
 * stores it (:1153-1156); 1: createQuaternionMsgFromRollPitchYaw(roll(), pitch(), yaw()) as the service does (:1255-1256). */
void scl_wire_loop_between(const float T_align[16], const float pose_cur6[6], const float pose_pre6[6], int quat_from_rpy, scl_transform* bet);

/* loop_info for an accepted intra-robot loop (:1143-1157) */
void scl_wire_make_loop_info(int robot, int index_cur, int index_pre, float fitness, const float T_icp[16], const float pose_cur6[6],
                             const float pose_pre6[6], scl_loop_info* out);

/* global_descriptor for keyframe `index` (:1004-1020): values = the engine's wire vector; has_pre = 0 leaves prePose zeroed (:1010) */
void scl_wire_make_global_descriptor(int index, const float* values, int n_values, const scl_transform* cur_pose, int has_pre,
                                     const float pre_pose6[6], scl_global_descriptor* out);

/* flat little-endian encoding of the three payloads (for transports that are not ROS): returns bytes written / needed
 * (buf may be NULL to size), decode returns bytes consumed or -1 on a short or inconsistent buffer */
int scl_wire_encode_global_descriptor(const scl_global_descriptor* m, unsigned char* buf, int cap);
int scl_wire_decode_global_descriptor(const unsigned char* buf, int len, scl_global_descriptor* m /* values points into buf */);
int scl_wire_encode_loop_info(const scl_loop_info* m, unsigned char* buf, int cap);
int scl_wire_decode_loop_info(const unsigned char* buf, int len, scl_loop_info* m);

#ifdef __cplusplus
}
#endif
#endif
