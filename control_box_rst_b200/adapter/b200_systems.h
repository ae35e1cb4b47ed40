// System dynamics of BASELINE.json configs[2] and configs[4] for the reference's object model.
//
// The reference ships no unicycle and no quadrotor (grep over /root/reference/src: 0 hits), so a user who wants those two benchmark
// OCPs behind corbo::StructuredOptimalControlProblem needs the model classes too.  They follow the pattern of the reference's own
// benchmark systems (src/systems/include/corbo-systems/benchmark/nonlinear_benchmark_systems.h:40-90): a SystemDynamicsInterface
// subclass with parameter getters and setters, registered in Factory<SystemDynamicsInterface> by class name, usable by ANY reference
// solver (the reference's LevenbergMarquardtSparse evaluates them on the host) -- and recognised by corbo::SolverB200Lm::describe(),
// which maps them to the device functors B200SQP_DYN_UNICYCLE / B200SQP_DYN_QUADROTOR (control_box_rst_b200/csrc/dynamics.cuh: the
// same equations in the same operation order; parameters travel through the getters, so nothing has to be repeated by hand).
#ifndef CONTROL_BOX_RST_B200_ADAPTER_B200_SYSTEMS_H_
#define CONTROL_BOX_RST_B200_ADAPTER_B200_SYSTEMS_H_

#include <corbo-systems/system_dynamics_interface.h>

#include <cassert>
#include <cmath>
#include <memory>

namespace corbo {

// Kinematic unicycle: x = [px, py, theta], u = [v, omega];  px' = v cos(theta), py' = v sin(theta), theta' = omega
class Unicycle : public SystemDynamicsInterface
{
 public:
    Unicycle() {}

    Ptr getInstance() const override { return std::make_shared<Unicycle>(); }
    bool isContinuousTime() const override { return true; }
    bool isLinear() const override { return false; }
    int getInputDimension() const override { return 2; }
    int getStateDimension() const override { return 3; }

    void dynamics(const Eigen::Ref<const StateVector>& x, const Eigen::Ref<const ControlVector>& u, Eigen::Ref<StateVector> f) const override
    {
        assert(x.size() == getStateDimension() && u.size() == getInputDimension() && f.size() == x.size());
        f[0] = u[0] * std::cos(x[2]);
        f[1] = u[0] * std::sin(x[2]);
        f[2] = u[1];
    }
};
FACTORY_REGISTER_SYSTEM_DYNAMICS(Unicycle)

// Rigid-body quadrotor, 12 states: x = [px py pz | phi theta psi | vx vy vz | p q r] (ZYX Euler angles, world-frame velocity, body
// rates), u = [T, tau_x, tau_y, tau_z] (total thrust, body torques); parameters mass, gravity, principal inertias
class Quadrotor : public SystemDynamicsInterface
{
 public:
    Quadrotor() {}
    Quadrotor(double mass, double gravity, double ixx, double iyy, double izz) : _m(mass), _g(gravity), _ixx(ixx), _iyy(iyy), _izz(izz) {}

    Ptr getInstance() const override { return std::make_shared<Quadrotor>(); }
    bool isContinuousTime() const override { return true; }
    bool isLinear() const override { return false; }
    int getInputDimension() const override { return 4; }
    int getStateDimension() const override { return 12; }

    void dynamics(const Eigen::Ref<const StateVector>& x, const Eigen::Ref<const ControlVector>& u, Eigen::Ref<StateVector> f) const override
    {
        assert(x.size() == getStateDimension() && u.size() == getInputDimension() && f.size() == x.size());
        const double sphi = std::sin(x[3]), cphi = std::cos(x[3]);
        const double sth = std::sin(x[4]), cth = std::cos(x[4]);
        const double spsi = std::sin(x[5]), cpsi = std::cos(x[5]);
        const double p = x[9], q = x[10], r = x[11];
        const double thrust_per_mass = u[0] / _m;
        f[0] = x[6];
        f[1] = x[7];
        f[2] = x[8];
        const double qr = q * sphi + r * cphi;
        f[3]  = p + qr * (sth / cth);
        f[4]  = q * cphi - r * sphi;
        f[5]  = qr / cth;
        f[6]  = (cphi * sth * cpsi + sphi * spsi) * thrust_per_mass;
        f[7]  = (cphi * sth * spsi - sphi * cpsi) * thrust_per_mass;
        f[8]  = cphi * cth * thrust_per_mass - _g;
        f[9]  = (u[1] + (_iyy - _izz) * q * r) / _ixx;
        f[10] = (u[2] + (_izz - _ixx) * p * r) / _iyy;
        f[11] = (u[3] + (_ixx - _iyy) * p * q) / _izz;
    }

    // access parameters
    void setParameters(double mass, double gravity, double ixx, double iyy, double izz)
    {
        _m   = mass;
        _g   = gravity;
        _ixx = ixx;
        _iyy = iyy;
        _izz = izz;
    }
    const double& getMass() const { return _m; }
    const double& getGravity() const { return _g; }
    const double& getInertiaXX() const { return _ixx; }
    const double& getInertiaYY() const { return _iyy; }
    const double& getInertiaZZ() const { return _izz; }

 private:
    double _m = 1.0, _g = 9.81, _ixx = 0.01, _iyy = 0.01, _izz = 0.02;
};
FACTORY_REGISTER_SYSTEM_DYNAMICS(Quadrotor)

}  // namespace corbo

#endif  // CONTROL_BOX_RST_B200_ADAPTER_B200_SYSTEMS_H_
