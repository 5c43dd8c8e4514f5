// ORACLE / TEST INFRASTRUCTURE ONLY -- not product code.
//
// Minimal stand-in for the part of Sophus (>= 1.0, header "sophus/geometry.hpp" as included at
// jac_Rt_gen_.cpp:9) that the reference's LM translation units use: SE3<double>::exp(.).matrix()
// (jac_Rt_gen_.cpp:419), SO3d::rotX/rotY/rotZ(.).matrix() and Constants<double>::pi()
// (sequence.hpp:14-21,40-45).  Written from Sophus' published algorithm: the rotation goes through
// the unit quaternion (cos(theta/2), sin(theta/2)/theta * omega) with Taylor factors below
// theta = 1e-10, the rotation matrix is the quaternion's, and the translation is V(omega) * upsilon
// with V = I + (1-cos)/theta^2 * hat(omega) + (theta-sin)/theta^3 * hat(omega)^2 (V = R below 1e-10).
// Tangent order is (upsilon, omega): translation first.
#ifndef EPIVO_ORACLE_SOPHUS_STANDIN
#define EPIVO_ORACLE_SOPHUS_STANDIN
#include <Eigen/Dense>
#include <cmath>

namespace Sophus {

template <class Scalar> struct Constants {
    static Scalar epsilon() { return Scalar(1e-10); }
    static Scalar pi() { return Scalar(3.141592653589793238462643383279502884); }
};

template <class Scalar> class SO3 {
    Scalar w_, x_, y_, z_;      // unit quaternion
  public:
    SO3() : w_(1), x_(0), y_(0), z_(0) {}
    static SO3 expAndTheta(Scalar ox, Scalar oy, Scalar oz, Scalar* theta) {
        const Scalar theta_sq = ox * ox + oy * oy + oz * oz;
        *theta = std::sqrt(theta_sq);
        const Scalar half_theta = Scalar(0.5) * (*theta);
        Scalar imag_factor, real_factor;
        if (*theta < Constants<Scalar>::epsilon()) {
            const Scalar theta_po4 = theta_sq * theta_sq;
            imag_factor = Scalar(0.5) - Scalar(1.0 / 48.0) * theta_sq + Scalar(1.0 / 3840.0) * theta_po4;
            real_factor = Scalar(1) - Scalar(0.5) * theta_sq + Scalar(1.0 / 384.0) * theta_po4;
        } else {
            const Scalar sin_half_theta = std::sin(half_theta);
            imag_factor = sin_half_theta / (*theta);
            real_factor = std::cos(half_theta);
        }
        SO3 q;
        q.w_ = real_factor; q.x_ = imag_factor * ox; q.y_ = imag_factor * oy; q.z_ = imag_factor * oz;
        return q;
    }
    static SO3 rotX(Scalar a) { Scalar th; return expAndTheta(a, 0, 0, &th); }
    static SO3 rotY(Scalar a) { Scalar th; return expAndTheta(0, a, 0, &th); }
    static SO3 rotZ(Scalar a) { Scalar th; return expAndTheta(0, 0, a, &th); }
    Eigen::MatrixXd matrix() const {
        const Scalar tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
        const Scalar twx = tx * w_, twy = ty * w_, twz = tz * w_;
        const Scalar txx = tx * x_, txy = ty * x_, txz = tz * x_;
        const Scalar tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        Eigen::MatrixXd R(3, 3);
        R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
        R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
        R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
        return R;
    }
};
typedef SO3<double> SO3d;

template <class Scalar> class SE3 {
    Eigen::MatrixXd T_;
  public:
    SE3() : T_(Eigen::MatrixXd::Identity(4, 4)) {}
    // a = (upsilon, omega), 6 x 1
    static SE3 exp(const Eigen::MatrixXd& a) {
        assert(a.size() == 6);
        const Scalar ux = a(0), uy = a(1), uz = a(2), ox = a(3), oy = a(4), oz = a(5);
        Scalar theta;
        const SO3<Scalar> so3 = SO3<Scalar>::expAndTheta(ox, oy, oz, &theta);
        Eigen::MatrixXd Omega(3, 3);
        Omega << 0, -oz, oy,
                 oz, 0, -ox,
                 -oy, ox, 0;
        const Eigen::MatrixXd Omega_sq = Omega * Omega;
        const Eigen::MatrixXd R = so3.matrix();
        Eigen::MatrixXd V(3, 3);
        if (theta < Constants<Scalar>::epsilon()) {
            V = R;
        } else {
            const Scalar theta_sq = theta * theta;
            V = Eigen::MatrixXd::Identity(3, 3) + (Scalar(1) - std::cos(theta)) / theta_sq * Omega +
                (theta - std::sin(theta)) / (theta_sq * theta) * Omega_sq;
        }
        Eigen::MatrixXd u(3, 1);
        u << ux, uy, uz;
        SE3 r;
        r.T_.template block<3, 3>(0, 0) = R;
        r.T_.template block<3, 1>(0, 3) = V * u;
        return r;
    }
    Eigen::MatrixXd matrix() const { return T_; }
};
typedef SE3<double> SE3d;

}  // namespace Sophus
#endif
