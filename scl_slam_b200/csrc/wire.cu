// wire.cu — host-side wire codec and pose arithmetic of the loop-closure path (include/scl_wire.h).
// No device code: the file is a .cu only so that the one nvcc recipe builds the whole library.
//
// Restated formulas (PCL / GTSAM / tf are not vendored in /root/reference; call sites: distributedMapping.h:1004-1024,
// 1129-1158, 1244-1259):
//   pcl::getTransformation(x,y,z,roll,pitch,yaw)      common/impl/eigen.hpp, Scalar = float
//   pcl::getTranslationAndEulerAngles(t, ...)         roll = atan2(t(2,1), t(2,2)); pitch = asin(-t(2,0)); yaw = atan2(t(1,0), t(0,0))
//   gtsam::Rot3::RzRyRx(x, y, z)                       Rz(z) * Ry(y) * Rx(x), double
//   gtsam::Pose3::between                              (R1^T R2, R1^T (t2 - t1))
//   gtsam::Rot3::toQuaternion                          Eigen::Quaterniond(matrix): the trace / largest-diagonal branches
//   gtsam::Rot3::roll/pitch/yaw                        xyz() by RQ decomposition
//   tf::createQuaternionMsgFromRollPitchYaw            tf::Quaternion::setRPY, double
#include "../../include/scl_wire.h"

#include <cmath>
#include <cstring>

namespace {

struct M3 { double m[3][3]; };

void pcl_transformation(const float p[6], float t[12])       /* row-major 3x4 */
{
    const float A = cosf(p[5]), B = sinf(p[5]), C = cosf(p[4]), D = sinf(p[4]), E = cosf(p[3]), F = sinf(p[3]), DE = D * E, DF = D * F;
    t[0] = A * C; t[1] = A * DF - B * E; t[2] = B * F + A * DE; t[3] = p[0];
    t[4] = B * C; t[5] = A * E + B * DF; t[6] = B * DE - A * F; t[7] = p[1];
    t[8] = -D;    t[9] = C * F;          t[10] = C * E;         t[11] = p[2];
}

M3 rzryrx(double x, double y, double z)
{
    const double cx = cos(x), sx = sin(x), cy = cos(y), sy = sin(y), cz = cos(z), sz = sin(z);
    const double ss_ = sx * sy, cs_ = cx * sy, sc_ = sx * cy, cc_ = cx * cy;
    const double c_s = cx * sz, s_s = sx * sz, _cs = cy * sz, _cc = cy * cz, s_c = sx * cz, c_c = cx * cz;
    const double ssc = ss_ * cz, csc = cs_ * cz, sss = ss_ * sz, css = cs_ * sz;
    M3 r = {{{_cc, -c_s + ssc, s_s + csc}, {_cs, c_c + sss, -s_c + css}, {-sy, sc_, cc_}}};
    return r;
}

void quat_from_matrix(const M3& r, double q[4] /* x y z w */)
{
    const double (*m)[3] = r.m;
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (m[2][1] - m[1][2]) * t; q[1] = (m[0][2] - m[2][0]) * t; q[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m[k][j] - m[j][k]) * t; q[j] = (m[j][i] + m[i][j]) * t; q[k] = (m[k][i] + m[i][k]) * t;
    }
}

void quat_from_rpy(double roll, double pitch, double yaw, double q[4])
{
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = cos(hy), sy = sin(hy), cp = cos(hp), sp = sin(hp), cr = cos(hr), sr = sin(hr);
    q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy; q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
}

M3 mul(const M3& a, const M3& b)
{
    M3 c;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}

// gtsam::RQ: R = Qz^T Qy^T Qx^T ... returns the xyz angles (roll, pitch, yaw) of a rotation matrix
void rot_xyz(const M3& A, double xyz[3])
{
    const double x = -atan2(-A.m[2][1], A.m[2][2]);
    const double cx = cos(-x), sx = sin(-x);
    const M3 Qx = {{{1, 0, 0}, {0, cx, -sx}, {0, sx, cx}}};
    const M3 B = mul(A, Qx);
    const double y = -atan2(B.m[2][0], B.m[2][2]);
    const double cy = cos(-y), sy = sin(-y);
    const M3 Qy = {{{cy, 0, sy}, {0, 1, 0}, {-sy, 0, cy}}};
    const M3 Cm = mul(B, Qy);
    const double z = -atan2(-Cm.m[1][0], Cm.m[1][1]);
    xyz[0] = x; xyz[1] = y; xyz[2] = z;
}

void put(unsigned char*& p, const void* v, int n) { memcpy(p, v, n); p += n; }
void get(const unsigned char*& p, void* v, int n) { memcpy(v, p, n); p += n; }

} // namespace

extern "C" {

void scl_wire_pose6_to_transform(const float pose6[6], scl_transform* out)
{
    double q[4];
    quat_from_rpy(pose6[3], pose6[4], pose6[5], q);
    out->tx = pose6[0]; out->ty = pose6[1]; out->tz = pose6[2];
    out->qx = q[0]; out->qy = q[1]; out->qz = q[2]; out->qw = q[3];
}

void scl_wire_loop_between(const float T_align[16], const float pose_cur6[6], const float pose_pre6[6], int quat_from_rpy_, scl_transform* bet)
{
    float w[12];
    pcl_transformation(pose_cur6, w);
    /* tfCorrect = T_align * tfWrong: Eigen's float product, each coefficient a left-to-right sum over k (the last row of both is 0 0 0 1) */
    float c[12];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) c[4 * i + j] = T_align[4 * i] * w[j] + T_align[4 * i + 1] * w[4 + j] + T_align[4 * i + 2] * w[8 + j];
        c[4 * i + 3] = T_align[4 * i] * w[3] + T_align[4 * i + 1] * w[7] + T_align[4 * i + 2] * w[11] + T_align[4 * i + 3];
    }
    const float x = c[3], y = c[7], z = c[11];
    const float roll = atan2f(c[9], c[10]), pitch = asinf(-c[8]), yaw = atan2f(c[4], c[0]);
    const M3 Rf = rzryrx(roll, pitch, yaw);
    const M3 Rt = rzryrx((double)pose_pre6[3], (double)pose_pre6[4], (double)pose_pre6[5]);
    const double d[3] = {(double)pose_pre6[0] - (double)x, (double)pose_pre6[1] - (double)y, (double)pose_pre6[2] - (double)z};
    M3 Rb; double tb[3];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Rb.m[i][j] = Rf.m[0][i] * Rt.m[0][j] + Rf.m[1][i] * Rt.m[1][j] + Rf.m[2][i] * Rt.m[2][j];
        tb[i] = Rf.m[0][i] * d[0] + Rf.m[1][i] * d[1] + Rf.m[2][i] * d[2];
    }
    double q[4];
    if (quat_from_rpy_) { double a[3]; rot_xyz(Rb, a); quat_from_rpy(a[0], a[1], a[2], q); }
    else quat_from_matrix(Rb, q);
    bet->tx = tb[0]; bet->ty = tb[1]; bet->tz = tb[2];
    bet->qx = q[0]; bet->qy = q[1]; bet->qz = q[2]; bet->qw = q[3];
}

void scl_wire_make_loop_info(int robot, int index_cur, int index_pre, float fitness, const float T_icp[16], const float pose_cur6[6],
                             const float pose_pre6[6], scl_loop_info* out)
{
    out->robot0 = robot; out->robot1 = robot; out->index0 = index_cur; out->index1 = index_pre; out->noise = fitness;
    scl_wire_loop_between(T_icp, pose_cur6, pose_pre6, 0, &out->bet_pose);
}

void scl_wire_make_global_descriptor(int index, const float* values, int n_values, const scl_transform* cur_pose, int has_pre,
                                     const float pre_pose6[6], scl_global_descriptor* out)
{
    memset(out, 0, sizeof(*out));
    out->index = index; out->values = values; out->n_values = n_values;
    if (cur_pose) out->cur_pose = *cur_pose;
    if (has_pre && pre_pose6) scl_wire_pose6_to_transform(pre_pose6, &out->pre_pose);
}

int scl_wire_encode_global_descriptor(const scl_global_descriptor* m, unsigned char* buf, int cap)
{
    const int need = 4 + 2 * 56 + 4 + 4 * (m->n_values > 0 ? m->n_values : 0);
    if (!buf) return need;
    if (cap < need) return -1;
    unsigned char* p = buf;
    put(p, &m->index, 4); put(p, &m->pre_pose, 56); put(p, &m->cur_pose, 56); put(p, &m->n_values, 4);
    if (m->n_values > 0) put(p, m->values, 4 * m->n_values);
    return need;
}

int scl_wire_decode_global_descriptor(const unsigned char* buf, int len, scl_global_descriptor* m)
{
    if (len < 120) return -1;
    const unsigned char* p = buf;
    get(p, &m->index, 4); get(p, &m->pre_pose, 56); get(p, &m->cur_pose, 56); get(p, &m->n_values, 4);
    if (m->n_values < 0 || (long long)len < 120 + 4ll * m->n_values) return -1;
    m->values = reinterpret_cast<const float*>(p);
    return 120 + 4 * m->n_values;
}

int scl_wire_encode_loop_info(const scl_loop_info* m, unsigned char* buf, int cap)
{
    const int need = 16 + 4 + 56;
    if (!buf) return need;
    if (cap < need) return -1;
    unsigned char* p = buf;
    put(p, &m->robot0, 4); put(p, &m->robot1, 4); put(p, &m->index0, 4); put(p, &m->index1, 4); put(p, &m->noise, 4); put(p, &m->bet_pose, 56);
    return need;
}

int scl_wire_decode_loop_info(const unsigned char* buf, int len, scl_loop_info* m)
{
    if (len < 76) return -1;
    const unsigned char* p = buf;
    get(p, &m->robot0, 4); get(p, &m->robot1, 4); get(p, &m->index0, 4); get(p, &m->index1, 4); get(p, &m->noise, 4); get(p, &m->bet_pose, 56);
    return 76;
}

} // extern "C"
