// TEST INFRASTRUCTURE ONLY.  Two system models that the reference does not ship (grep for
// unicycle/quadrotor in /root/reference -> 0 hits) written against the reference's own corbo::SystemDynamicsInterface
// (src/systems/include/corbo-systems/system_dynamics_interface.h:66-121) so that BASELINE.json configs 3 and 5 can be run
// through the unmodified reference solver.  The same equations, in the same operation order, are restated in
// oracle/sqp_oracle.cpp and in control_box_rst_b200/csrc/dynamics.cuh.
#ifndef ORACLE_REF_MODELS_H_
#define ORACLE_REF_MODELS_H_

#include <corbo-systems/system_dynamics_interface.h>

#include <cmath>
#include <memory>

namespace b200ref {

// x = [px, py, theta], u = [v, omega]
class Unicycle : public corbo::SystemDynamicsInterface
{
 public:
    Ptr getInstance() const override { return std::make_shared<Unicycle>(); }
    bool isContinuousTime() const override { return true; }
    bool isLinear() const override { return false; }
    int getInputDimension() const override { return 2; }
    int getStateDimension() const override { return 3; }
    void dynamics(const Eigen::Ref<const StateVector>& x, const Eigen::Ref<const ControlVector>& u, Eigen::Ref<StateVector> f) const override
    {
        f[0] = u[0] * std::cos(x[2]);
        f[1] = u[0] * std::sin(x[2]);
        f[2] = u[1];
    }
};

// x = [px py pz | phi theta psi | vx vy vz | p q r], u = [T, tau_x, tau_y, tau_z]; ZYX Euler angles, world-frame velocity
class Quadrotor : public corbo::SystemDynamicsInterface
{
 public:
    Quadrotor() {}
    Quadrotor(double m, double g, double ixx, double iyy, double izz) : _m(m), _g(g), _ixx(ixx), _iyy(iyy), _izz(izz) {}
    Ptr getInstance() const override { return std::make_shared<Quadrotor>(); }
    bool isContinuousTime() const override { return true; }
    bool isLinear() const override { return false; }
    int getInputDimension() const override { return 4; }
    int getStateDimension() const override { return 12; }
    void dynamics(const Eigen::Ref<const StateVector>& x, const Eigen::Ref<const ControlVector>& u, Eigen::Ref<StateVector> f) const override
    {
        const double sphi = std::sin(x[3]), cphi = std::cos(x[3]);
        const double sth = std::sin(x[4]), cth = std::cos(x[4]);
        const double spsi = std::sin(x[5]), cpsi = std::cos(x[5]);
        const double p = x[9], q = x[10], r = x[11];
        const double tm = u[0] / _m;
        f[0] = x[6];
        f[1] = x[7];
        f[2] = x[8];
        const double qr = q * sphi + r * cphi;
        f[3] = p + qr * (sth / cth);
        f[4] = q * cphi - r * sphi;
        f[5] = qr / cth;
        f[6] = (cphi * sth * cpsi + sphi * spsi) * tm;
        f[7] = (cphi * sth * spsi - sphi * cpsi) * tm;
        f[8] = cphi * cth * tm - _g;
        f[9]  = (u[1] + (_iyy - _izz) * q * r) / _ixx;
        f[10] = (u[2] + (_izz - _ixx) * p * r) / _iyy;
        f[11] = (u[3] + (_ixx - _iyy) * p * q) / _izz;
    }

 private:
    double _m = 1.0, _g = 9.81, _ixx = 0.01, _iyy = 0.01, _izz = 0.02;
};

}  // namespace b200ref

#endif  // ORACLE_REF_MODELS_H_
