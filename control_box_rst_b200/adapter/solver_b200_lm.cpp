// corbo::SolverB200Lm: hypergraph walk -> b200sqp_ocp descriptor -> libb200sqp.so.  See solver_b200_lm.h.
#include "solver_b200_lm.h"

#include <corbo-core/console.h>
#include <corbo-numerics/explicit_integrators.h>
#include <corbo-optimal-control/functions/final_state_constraints.h>
#include <corbo-optimal-control/functions/final_state_cost.h>
#include <corbo-optimal-control/functions/minimum_time.h>
#include <corbo-optimal-control/functions/quadratic_cost.h>
#include <corbo-optimal-control/structured_ocp/edges/finite_differences_collocation_edges.h>
#include <corbo-optimal-control/structured_ocp/edges/misc_edges.h>
#include <corbo-optimal-control/structured_ocp/edges/multiple_shooting_edges.h>
#include <corbo-optimization/hyper_graph/hyper_graph_optimization_problem_base.h>
#include <corbo-systems/benchmark/linear_benchmark_systems.h>
#include <corbo-systems/benchmark/nonlinear_benchmark_systems.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>

#include "b200_systems.h"

namespace corbo {

// The per-object halves of a batched solve (reading x0 / the parameter vector out of the vertex objects, writing the result back) touch
// one reference object each and nothing shared: they run on all host threads for large batches (8.5 us per object on one thread
// would otherwise be 100x the device time of the whole batch).
template <class F>
static void forEachObject(int count, F&& body)
{
    const int hw = (int)std::thread::hardware_concurrency();
    const int threads = count >= 256 ? std::max(1, std::min(hw > 0 ? hw : 1, count / 64)) : 1;
    if (threads == 1)
    {
        for (int i = 0; i < count; ++i) body(i);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (int i = (int)((long long)count * t / threads), e = (int)((long long)count * (t + 1) / threads); i < e; ++i) body(i);
        });
    for (int i = 0, e = count / threads; i < e; ++i) body(i);
    for (auto& th : pool) th.join();
}

// Structure identity of two descriptors: everything but LIVE vertex values.  On a variable-dt grid dt_ref is read from the first dt
// vertex, i.e. it is the optimised value of the previous solve and differs per instance; the device only uses it to initialise
// trajectories, which the adapter never asks for (it uploads the parameters).  Fixed-dt grids keep it: there it IS the structure.
static bool sameStructure(const b200sqp_ocp& a, const b200sqp_ocp& b)
{
    b200sqp_ocp x = a, y = b;
    if (x.grid == B200SQP_GRID_FD_NONUNIFORM_VARDT) x.dt_ref = 0.0;
    if (y.grid == B200SQP_GRID_FD_NONUNIFORM_VARDT) y.dt_ref = 0.0;
    return std::memcmp(&x, &y, sizeof(x)) == 0;
}

SolverB200Lm::SolverB200Lm()
{
    // LevenbergMarquardtSparse defaults (levenberg_marquardt_sparse.h:112-124)
    _opts = {10, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    std::memset(&_ocp, 0, sizeof(_ocp));
    std::memset(&_dims, 0, sizeof(_dims));
}

SolverB200Lm::~SolverB200Lm() { clear(); }

void SolverB200Lm::setPenaltyWeights(double weight_eq, double weight_ineq, double weight_bounds)
{
    _opts.weight_eq     = weight_eq;
    _opts.weight_ineq   = weight_ineq;
    _opts.weight_bounds = weight_bounds;
}

void SolverB200Lm::setWeightAdapation(double factor_eq, double factor_ineq, double factor_bounds, double max_eq, double max_ineq, double max_bounds)
{
    _opts.adapt_factor_eq     = factor_eq;
    _opts.adapt_factor_ineq   = factor_ineq;
    _opts.adapt_factor_bounds = factor_bounds;
    _opts.adapt_max_eq        = max_eq;
    _opts.adapt_max_ineq      = max_ineq;
    _opts.adapt_max_bounds    = max_bounds;
}

SolverStatus SolverB200Lm::fail(const std::string& msg)
{
    _error = msg;
    PRINT_ERROR("SolverB200Lm: " << msg);
    return SolverStatus::Error;
}

bool SolverB200Lm::initialize(OptimizationProblemInterface* problem)
{
    // same contract as LevenbergMarquardtSparse::initialize (levenberg_marquardt_sparse.cpp:32-41)
    if (problem && !problem->isLeastSquaresProblem())
    {
        PRINT_ERROR("SolverB200Lm(): cannot handle non-least-squares objectives or LS objectives in non-LS form.");
        return false;
    }
    if (!b200sqp_device_available())
    {
        PRINT_ERROR("SolverB200Lm(): no sm_100 class CUDA device visible; this solver has no CPU path.");
        return false;
    }
    return true;
}

void SolverB200Lm::clear()
{
    if (_handle) b200sqp_destroy(_handle);
    _handle = nullptr;
    _batch  = 0;
}

double SolverB200Lm::lastSolveMilliseconds() const
{
    float ms = 0;
    if (_handle && b200sqp_last_solve_ms(_handle, &ms) == 0) return ms;
    return -1;
}

// Walk the hypergraph (BaseHyperGraphOptimizationProblem::getGraph, hyper_graph_optimization_problem_base.h:89) and the functor
// objects into the flat descriptor of include/b200sqp.h.
bool SolverB200Lm::describe(OptimizationProblemInterface& problem, b200sqp_ocp& d, std::vector<double>& x0, std::vector<double>& xref)
{
    auto* hg = dynamic_cast<BaseHyperGraphOptimizationProblem*>(&problem);
    if (!hg || !hg->getGraph().hasEdgeSet())
    {
        _error = "problem is not a hypergraph optimization problem";
        return false;
    }
    OptimizationEdgeSet* edges = hg->getGraph().getEdgeSetRaw();
    if (!edges->getMixedEdgesRef().empty() || !edges->getObjectiveEdgesRef().empty())
    {
        _error = "mixed / non-lsq objective edges are outside the device registry";
        return false;
    }
    // final-stage constraint (functions/final_state_constraints.h): one extra equality edge (TerminalEqualityConstraint) or the
    // only inequality edge (TerminalBall, diagonal S), a unary edge on xf created after all interval edges
    // (finite_differences_grid.cpp:131-144)
    auto* term_eq   = dynamic_cast<TerminalEqualityConstraint*>(_final_constraint.get());
    auto* term_ball = dynamic_cast<TerminalBall*>(_final_constraint.get());
    if (_final_constraint && !term_eq && !term_ball)
    {
        _error = "final-stage constraint type is not in the device registry";
        return false;
    }
    std::vector<BaseEdge::Ptr> eq = edges->getEqualityEdgesRef();  // copy of the pointer list: the final-stage edge is split off below
    std::vector<BaseEdge::Ptr>& ineq = edges->getInequalityEdgesRef();
    bool has_term_eq = false, has_term_ball = false;
    if (term_eq && !eq.empty() && eq.back()->getNumVertices() == 1)
    {
        has_term_eq = true;
        eq.pop_back();
    }
    if (!ineq.empty())
    {
        if (!(term_ball && ineq.size() == 1 && ineq.front()->getNumVertices() == 1 && ineq.front()->getDimension() == 1))
        {
            _error = "inequality edges other than one TerminalBall edge are outside the device registry";
            return false;
        }
        has_term_ball = true;
    }
    // TwoScalarEqualEdges (NonUniformFiniteDifferencesVariableGrid::setDtEqConstraint, non_uniform_finite_differences_variable_grid.cpp:150-154)
    // sit between the dynamics edges, one after the dynamics edge of every interval k >= 1: split them off, the descriptor carries a flag
    int n_dt_eq = 0;
    {
        std::vector<BaseEdge::Ptr> dyn_edges;
        for (size_t i = 0; i < eq.size(); ++i)
        {
            if (dynamic_cast<TwoScalarEqualEdge*>(eq[i].get()))
            {
                // expected position: right after the dynamics edge of interval k = number of dynamics edges so far - 1 >= 1
                if (dyn_edges.size() < 2 || (int)dyn_edges.size() - 1 != n_dt_eq + 1)
                {
                    _error = "TwoScalarEqualEdge at an unexpected position";
                    return false;
                }
                ++n_dt_eq;
            }
            else
                dyn_edges.push_back(eq[i]);
        }
        eq.swap(dyn_edges);
    }
    if (eq.empty() || !_dynamics)
    {
        _error = "no dynamics edges, or setSystemDynamics() was not called";
        return false;
    }
    if (n_dt_eq != 0 && n_dt_eq != (int)eq.size() - 1)
    {
        _error = "dt equality edges are expected between all consecutive intervals or none";
        return false;
    }
    std::memset(&d, 0, sizeof(d));
    d.dt_eq_constraint = n_dt_eq > 0 ? 1 : 0;
    const int K = (int)eq.size();
    d.n_grid    = K + 1;
    d.nx        = _dynamics->getStateDimension();
    d.nu        = _dynamics->getInputDimension();
    if (d.nx > B200SQP_MAX_NX || d.nu > B200SQP_MAX_NU)
    {
        _error = "state/input dimension exceeds the device limits";
        return false;
    }

    // ---- dynamics registry
    if (auto* s = dynamic_cast<VanDerPolOscillator*>(_dynamics.get()))
    {
        d.dynamics      = B200SQP_DYN_VAN_DER_POL;
        d.dyn_params[0] = s->getDampingCoefficient();
    }
    else if (dynamic_cast<CartPole*>(_dynamics.get()))
        d.dynamics = B200SQP_DYN_CART_POLE;  // parameters are private constants in the reference
    else if (auto* s = dynamic_cast<SerialIntegratorSystem*>(_dynamics.get()))
    {
        if (s->getDimension() < 2 || s->getDimension() > 4)
        {
            _error = "SerialIntegratorSystem: only dimensions 2, 3 and 4 are in the device registry";
            return false;
        }
        d.dynamics = s->getDimension() == 2 ? B200SQP_DYN_DOUBLE_INTEGRATOR : s->getDimension() == 3 ? B200SQP_DYN_TRIPLE_INTEGRATOR : B200SQP_DYN_QUAD_INTEGRATOR;
        d.dyn_params[0] = s->getTimeConstant();
    }
    else if (dynamic_cast<DuffingOscillator*>(_dynamics.get()) || dynamic_cast<SimplePendulum*>(_dynamics.get()) ||
             dynamic_cast<MasslessPendulum*>(_dynamics.get()) || dynamic_cast<ToyExample*>(_dynamics.get()) ||
             (dynamic_cast<LinearStateSpaceModel*>(_dynamics.get()) &&
              ((d.nx == 2 && d.nu == 1) || (d.nx == 3 && d.nu == 1) || (d.nx == 4 && d.nu == 1) || (d.nx == 4 && d.nu == 2))))
    {
        // setters without getters in the reference: the user repeats the values through setSystemDynamicsParameters(); selfCheck()
        // compares the device residuals with the reference's own computeValues after the upload, so a wrong value cannot go unnoticed
        int count = 0;
        if (dynamic_cast<DuffingOscillator*>(_dynamics.get())) { d.dynamics = B200SQP_DYN_DUFFING; count = 3; }
        else if (dynamic_cast<SimplePendulum*>(_dynamics.get())) { d.dynamics = B200SQP_DYN_SIMPLE_PENDULUM; count = 4; }
        else if (dynamic_cast<MasslessPendulum*>(_dynamics.get())) { d.dynamics = B200SQP_DYN_MASSLESS_PENDULUM; count = 1; }
        else if (dynamic_cast<ToyExample*>(_dynamics.get())) { d.dynamics = B200SQP_DYN_TOY_EXAMPLE; count = 1; }
        else  // LinearStateSpaceModel: A (column-major) then B (column-major), as given to setParameters(A, B)
        {
            d.dynamics = d.nx == 2 ? B200SQP_DYN_LINEAR_2X1 : d.nx == 3 ? B200SQP_DYN_LINEAR_3X1 : d.nu == 1 ? B200SQP_DYN_LINEAR_4X1 : B200SQP_DYN_LINEAR_4X2;
            count      = d.nx * d.nx + d.nx * d.nu;
        }
        if ((int)_dynamics_parameters.size() != count)
        {
            _error = "this system dynamics class exposes no parameter getters: hand its " + std::to_string(count) +
                     " parameter(s) to setSystemDynamicsParameters()";
            return false;
        }
        for (int i = 0; i < count; ++i) d.dyn_params[i] = _dynamics_parameters[i];
    }
    else if (dynamic_cast<Unicycle*>(_dynamics.get()))
        d.dynamics = B200SQP_DYN_UNICYCLE;  // b200_systems.h; no parameters
    else if (auto* s = dynamic_cast<Quadrotor*>(_dynamics.get()))
    {
        d.dynamics      = B200SQP_DYN_QUADROTOR;  // b200_systems.h; parameters through its getters
        d.dyn_params[0] = s->getMass();
        d.dyn_params[1] = s->getGravity();
        d.dyn_params[2] = s->getInertiaXX();
        d.dyn_params[3] = s->getInertiaYY();
        d.dyn_params[4] = s->getInertiaZZ();
    }
    else if (dynamic_cast<FreeSpaceRocket*>(_dynamics.get()))
        d.dynamics = B200SQP_DYN_FREE_SPACE_ROCKET;  // no parameters
    else if (dynamic_cast<ArtsteinsCircle*>(_dynamics.get()))
        d.dynamics = B200SQP_DYN_ARTSTEINS_CIRCLE;  // no parameters
    else
    {
        _error = "system dynamics type is not in the device registry";
        return false;
    }

    // ---- grid kind from the equality edges: every edge must be a dynamics edge over (x_k, u_k, x_{k+1}, dt_k)
    bool fd = true, ms = true;
    for (BaseEdge::Ptr& e : eq)
    {
        fd = fd && dynamic_cast<FDCollocationEdge*>(e.get()) != nullptr;
        ms = ms && dynamic_cast<MSVariableDynamicsOnlyEdge*>(e.get()) != nullptr;
        if (e->getNumVertices() != 4 || e->getDimension() != d.nx)
        {
            _error = "unexpected equality edge shape";
            return false;
        }
    }
    if (!fd && !ms)
    {
        _error = "equality edges are neither FDCollocationEdge nor MSVariableDynamicsOnlyEdge";
        return false;
    }
    const VertexInterface* dt0 = eq.front()->getVertexRaw(3);
    bool single_dt             = true;
    for (BaseEdge::Ptr& e : eq) single_dt = single_dt && e->getVertexRaw(3) == dt0;
    if (single_dt && !dt0->isFixed())
    {
        _error = "a single free dt (arrow-structured Hessian) is not supported yet";
        return false;
    }
    if (ms)
        d.grid = B200SQP_GRID_MULTIPLE_SHOOTING;
    else
        d.grid = single_dt ? B200SQP_GRID_FD_UNIFORM : B200SQP_GRID_FD_NONUNIFORM_VARDT;
    d.dt_ref = dt0->getData()[0];
    // variable-dt grids: a live vertex value, not structure (sameStructure() masks it); it only has to be a valid step size
    if (d.grid == B200SQP_GRID_FD_NONUNIFORM_VARDT && !(d.dt_ref > 0)) d.dt_ref = 0.1;
    d.dt_lb  = dt0->getLowerBounds()[0];
    d.dt_ub  = dt0->getUpperBounds()[0];
    if (ms)
    {
        if (dynamic_cast<IntegratorExplicitRungeKutta4*>(_integrator.get()))
            d.integrator = B200SQP_INT_RK4;
        else if (dynamic_cast<IntegratorExplicitEuler*>(_integrator.get()))
            d.integrator = B200SQP_INT_EULER;
        else
        {
            _error = "setIntegrator(): Euler or RK4 expected";
            return false;
        }
    }
    else
    {
        // grids default to Crank-Nicolson (full_discretization_grid_base.h:138)
        const FiniteDifferencesCollocationInterface* c = _collocation.get();
        if (!c || dynamic_cast<const CrankNicolsonDiffCollocation*>(c))
            d.collocation = B200SQP_COLL_CRANK_NICOLSON;
        else if (dynamic_cast<const ForwardDiffCollocation*>(c))
            d.collocation = B200SQP_COLL_FORWARD;
        else if (dynamic_cast<const BackwardDiffCollocation*>(c))
            d.collocation = B200SQP_COLL_BACKWARD;
        else if (dynamic_cast<const MidpointDiffCollocation*>(c))
            d.collocation = B200SQP_COLL_MIDPOINT;
        else
        {
            _error = "collocation type is not in the device registry";
            return false;
        }
    }

    // ---- vertices: x0 (must be fixed), bounds, fixed goal mask
    const VertexInterface* x_first = eq.front()->getVertexRaw(0);
    const VertexInterface* u_first = eq.front()->getVertexRaw(1);
    const VertexInterface* x_last  = eq.back()->getVertexRaw(2);
    if (!x_first->isFixed())
    {
        _error = "the first state vertex is expected to be fixed";
        return false;
    }
    x0.assign(x_first->getData(), x_first->getData() + d.nx);
    for (int i = 0; i < d.nx; ++i)
    {
        d.x_lb[i]     = x_last->getLowerBounds()[i];
        d.x_ub[i]     = x_last->getUpperBounds()[i];
        d.xf_fixed[i] = x_last->isFixedComponent(i) ? 1 : 0;
    }
    for (int i = 0; i < d.nu; ++i)
    {
        d.u_lb[i] = u_first->getLowerBounds()[i];
        d.u_ub[i] = u_first->getUpperBounds()[i];
    }

    // ---- reference: static state reference (also the value of fixed goal components)
    xref.assign(d.nx, 0.0);
    if (_xref)
    {
        // a static reference is one vector; of a non-static one (reference_trajectory.h:60-95) the LAST cached row is the "static" part
        // (goal, fixed goal components, final-stage constraint) and all rows go to the device through b200sqp_set_reference_trajectory
        const ReferenceTrajectoryInterface::OutputVector& r = _xref->getReferenceCached(_xref->isStatic() ? 0 : K);
        for (int i = 0; i < d.nx; ++i) xref[i] = r[i];
    }
    else
        for (int i = 0; i < d.nx; ++i)
            if (d.xf_fixed[i]) xref[i] = x_last->getData()[i];
    d.zero_u_ref = 1;
    d.zero_x_ref = 1;
    for (double v : xref) d.zero_x_ref = d.zero_x_ref && v == 0.0;

    // ---- costs
    d.stage_cost = B200SQP_COST_NONE;
    if (_stage_cost)
    {
        if (auto* q = dynamic_cast<QuadraticFormCost*>(_stage_cost.get()))
        {
            if (!q->isLsqFormNonIntegralStateTerm(0) || q->hasIntegralTerms(0))
            {
                _error = "QuadraticFormCost must be non-integral and in lsq form";
                return false;
            }
            const Eigen::MatrixXd& Q = q->getWeightQ();
            const Eigen::MatrixXd& R = q->getWeightR();
            if (Q.rows() != d.nx || Q.cols() != d.nx || R.rows() != d.nu || R.cols() != d.nu)
            {
                _error = "QuadraticFormCost: Q (nx x nx) and R (nu x nu) expected";
                return false;
            }
            d.stage_cost = B200SQP_COST_QUADRATIC_LSQ;
            for (int i = 0; i < d.nx; ++i) d.q_diag[i] = Q(i, i);
            for (int i = 0; i < d.nu; ++i) d.r_diag[i] = R(i, i);
            // full matrices travel as they are: the library takes their square roots like setWeightQ / setWeightR do (diagonal to 1e-10 ->
            // element-wise, else the upper Cholesky factor; quadratic_cost.cpp:32-96)
            if (!Q.isDiagonal(1e-10))
            {
                d.q_dense = 1;
                for (int i = 0; i < d.nx; ++i)
                    for (int j = 0; j < d.nx; ++j) d.q_full[i * d.nx + j] = Q(i, j);
            }
            if (!R.isDiagonal(1e-10))
            {
                d.r_dense = 1;
                for (int i = 0; i < d.nu; ++i)
                    for (int j = 0; j < d.nu; ++j) d.r_full[i * d.nu + j] = R(i, j);
            }
        }
        else if (auto* t = dynamic_cast<MinimumTime*>(_stage_cost.get()))
        {
            if (!t->isLsqFormNonIntegralDtTerm(0))
            {
                _error = "MinimumTime must be in lsq form";
                return false;
            }
            d.stage_cost = B200SQP_COST_MINIMUM_TIME_LSQ;
        }
        else
        {
            _error = "stage cost type is not in the device registry";
            return false;
        }
    }
    d.final_cost = 0;
    if (_final_cost)
    {
        auto* qf = dynamic_cast<QuadraticFinalStateCost*>(_final_cost.get());
        if (!qf || !qf->isLsqFormNonIntegralStateTerm(0) || qf->getWeightQf().rows() != d.nx || qf->getWeightQf().cols() != d.nx)
        {
            _error = "final cost must be a QuadraticFinalStateCost (nx x nx) in lsq form";
            return false;
        }
        d.final_cost = 1;
        for (int i = 0; i < d.nx; ++i) d.qf_diag[i] = qf->getWeightQf()(i, i);
        if (!qf->getWeightQf().isDiagonal(1e-10))
        {
            d.qf_dense = 1;
            for (int i = 0; i < d.nx; ++i)
                for (int j = 0; j < d.nx; ++j) d.qf_full[i * d.nx + j] = qf->getWeightQf()(i, j);
        }
    }
    d.final_constraint = B200SQP_FINAL_CONSTRAINT_NONE;
    if (has_term_eq)
    {
        if (term_eq->getXRef().size() != d.nx)
        {
            _error = "TerminalEqualityConstraint: xref dimension does not match the state dimension";
            return false;
        }
        d.final_constraint = B200SQP_FINAL_CONSTRAINT_EQUALITY;
        for (int i = 0; i < d.nx; ++i) d.term_xref[i] = term_eq->getXRef()[i];
    }
    else if (has_term_ball)
    {
        // TerminalBallInheritFromCost (final_state_constraints.h:98-128): the same row with S = Qf of the quadratic final cost
        // (its update() copies the weight, final_state_constraints.cpp:170-194) and its OWN gamma member, which shadows TerminalBall's
        auto* inherit = dynamic_cast<TerminalBallInheritFromCost*>(_final_constraint.get());
        Eigen::MatrixXd S = term_ball->getWeightS();
        double gamma      = term_ball->getGamma();
        if (inherit)
        {
            auto* qf = dynamic_cast<QuadraticFinalStateCost*>(_final_cost.get());
            if (!qf)
            {
                _error = "TerminalBallInheritFromCost needs the QuadraticFinalStateCost handed to setFinalStageCost()";
                return false;
            }
            S     = qf->getWeightQf();
            gamma = inherit->_gamma;
        }
        if (S.rows() != d.nx || !S.isDiagonal(1e-10))  // TerminalBall::setWeightS switches to its diagonal mode by the same test
        {
            _error = "TerminalBall: only a diagonal weight S is in the device registry";
            return false;
        }
        d.final_constraint = B200SQP_FINAL_CONSTRAINT_BALL;
        for (int i = 0; i < d.nx; ++i) d.term_s_diag[i] = S(i, i);
        d.term_gamma = gamma;
    }
    return true;
}

bool SolverB200Lm::upload(OptimizationProblemInterface& problem, int batch, std::vector<double>* x0_out, std::vector<double>* xref_out)
{
    std::vector<double> x0, xref;
    b200sqp_ocp d;
    if (!describe(problem, d, x0, xref)) return false;
    b200sqp_dims dims;
    if (b200sqp_dims_of(&d, &dims) != 0)
    {
        _error = std::string("descriptor rejected: ") + b200sqp_last_error();
        return false;
    }
    // the reference's own numbers (levenberg_marquardt_sparse.cpp:56-71) must agree with the device structure, index for index
    if (dims.n_params != problem.getParameterDimension() || dims.m_lsq != problem.getLsqObjectiveDimension() ||
        dims.m_eq != problem.getEqualityDimension() || dims.m_ineq != problem.getInequalityDimension() ||
        dims.m_bounds != problem.finiteCombinedBoundsDimension())
    {
        _error = "device structure does not match the hypergraph's dimensions";
        return false;
    }
    if (_handle && (!sameStructure(d, _ocp) || batch != _batch)) clear();
    _fresh = false;
    if (!_handle)
    {
        _fresh = true;
        if (b200sqp_create(&d, batch, _device, &_handle) != 0)
        {
            _error  = std::string("b200sqp_create: ") + b200sqp_last_error();
            _handle = nullptr;
            return false;
        }
        _ocp   = d;
        _dims  = dims;
        _batch = batch;
    }
    if (x0_out) *x0_out = x0;
    if (xref_out) *xref_out = xref;
    return true;
}

// Per-instance data of a further problem of a batch whose structure `_ocp` was derived from the first one: start state and
// reference, plus the cheap structural checks (dimensions, edge counts, bounds) -- not another full hypergraph walk per object.
bool SolverB200Lm::instanceData(OptimizationProblemInterface& problem, double* x0, double* xref, std::string& error) const
{
    auto* hg = dynamic_cast<BaseHyperGraphOptimizationProblem*>(&problem);
    if (!hg || !hg->getGraph().hasEdgeSet())
    {
        error = "problem is not a hypergraph optimization problem";
        return false;
    }
    if (problem.getParameterDimension() != _dims.n_params || problem.getLsqObjectiveDimension() != _dims.m_lsq ||
        problem.getEqualityDimension() != _dims.m_eq || problem.getInequalityDimension() != _dims.m_ineq ||
        problem.finiteCombinedBoundsDimension() != _dims.m_bounds)
    {
        error = "dimensions differ from the first problem of the batch";
        return false;
    }
    OptimizationEdgeSet* edges        = hg->getGraph().getEdgeSetRaw();
    std::vector<BaseEdge::Ptr>& eq    = edges->getEqualityEdgesRef();
    const int K                       = _ocp.n_grid - 1;
    const int n_dt_eq                 = (_ocp.dt_eq_constraint && K > 1) ? K - 1 : 0;
    const int n_eq                    = K + n_dt_eq + (_ocp.final_constraint == B200SQP_FINAL_CONSTRAINT_EQUALITY ? 1 : 0);
    // the last dynamics edge: dynamics edges and (from interval 1 on) dt equality edges alternate
    const int last_dyn                = n_dt_eq ? 2 * K - 3 : K - 1;
    if ((int)eq.size() != n_eq || !edges->getMixedEdgesRef().empty() || !edges->getObjectiveEdgesRef().empty() ||
        eq.front()->getNumVertices() != 4 || eq[last_dyn]->getNumVertices() != 4)
    {
        error = "edge lists differ from the first problem of the batch";
        return false;
    }
    const VertexInterface* x_first = eq.front()->getVertexRaw(0);
    const VertexInterface* u_first = eq.front()->getVertexRaw(1);
    const VertexInterface* dt0     = eq.front()->getVertexRaw(3);
    const VertexInterface* x_last  = eq[last_dyn]->getVertexRaw(2);
    bool same = x_first->isFixed() && x_first->getDimension() == _ocp.nx && u_first->getDimension() == _ocp.nu;
    for (int i = 0; same && i < _ocp.nx; ++i)
        same = x_last->getLowerBounds()[i] == _ocp.x_lb[i] && x_last->getUpperBounds()[i] == _ocp.x_ub[i] &&
               (x_last->isFixedComponent(i) ? 1 : 0) == _ocp.xf_fixed[i];
    for (int i = 0; same && i < _ocp.nu; ++i) same = u_first->getLowerBounds()[i] == _ocp.u_lb[i] && u_first->getUpperBounds()[i] == _ocp.u_ub[i];
    same = same && dt0->getLowerBounds()[0] == _ocp.dt_lb && dt0->getUpperBounds()[0] == _ocp.dt_ub;
    if (_ocp.grid != B200SQP_GRID_FD_NONUNIFORM_VARDT) same = same && dt0->getData()[0] == _ocp.dt_ref;
    if (!same)
    {
        error = "bounds, fixed components or step size differ from the first problem of the batch";
        return false;
    }
    for (int i = 0; i < _ocp.nx; ++i) x0[i] = x_first->getData()[i];
    if (_xref)
    {
        const ReferenceTrajectoryInterface::OutputVector& r = _xref->getReferenceCached(_xref->isStatic() ? 0 : K);
        for (int i = 0; i < _ocp.nx; ++i) xref[i] = r[i];
    }
    else
        for (int i = 0; i < _ocp.nx; ++i) xref[i] = _ocp.xf_fixed[i] ? x_last->getData()[i] : 0.0;
    return true;
}

// A non-static state reference (setStateReference): its cached rows 0..N-1 -- the same for every problem of a batch, the reference object
// is the solver's -- go to the device as a time-varying reference.  Static references need nothing here.
bool SolverB200Lm::uploadReferenceTrajectory(int batch)
{
    if (!_xref || _xref->isStatic()) return true;
    const int N = _ocp.n_grid, nx = _ocp.nx;
    std::vector<double> rows((size_t)batch * N * nx);
    for (int k = 0; k < N; ++k)
    {
        const ReferenceTrajectoryInterface::OutputVector& r = _xref->getReferenceCached(k);
        for (int i = 0; i < nx; ++i) rows[(size_t)k * nx + i] = r[i];
    }
    for (int b = 1; b < batch; ++b) std::copy(rows.begin(), rows.begin() + (size_t)N * nx, rows.begin() + (size_t)b * N * nx);
    return b200sqp_set_reference_trajectory(_handle, rows.data()) == 0;
}

// Guard of SURVEY.md section 8b: the device residual vector at the current parameters must equal the reference's computeValues.
bool SolverB200Lm::selfCheck(OptimizationProblemInterface& problem)
{
    const int m = _dims.m_lsq + _dims.m_eq + _dims.m_ineq + _dims.m_bounds;
    std::vector<double> dev((size_t)m * _batch);
    if (b200sqp_evaluate(_handle, 1.0, 1.0, 1.0, dev.data(), nullptr) != 0)
    {
        _error = std::string("b200sqp_evaluate: ") + b200sqp_last_error();
        return false;
    }
    Eigen::VectorXd host(m);
    if (_dims.m_lsq > 0) problem.computeValuesLsqObjective(host.segment(0, _dims.m_lsq));
    if (_dims.m_eq > 0) problem.computeValuesEquality(host.segment(_dims.m_lsq, _dims.m_eq));
    if (_dims.m_ineq > 0) problem.computeValuesActiveInequality(host.segment(_dims.m_lsq + _dims.m_eq, _dims.m_ineq), 1.0);
    if (_dims.m_bounds > 0)
        problem.computeDistanceFiniteCombinedBounds(host.segment(_dims.m_lsq + _dims.m_eq + _dims.m_ineq, _dims.m_bounds));
    double worst = 0;
    for (int i = 0; i < m; ++i) worst = std::max(worst, std::abs(host[i] - dev[i]) / std::max(1.0, std::abs(host[i])));
    if (!(worst <= 1e-9))
    {
        _error = "device residuals differ from the reference's computeValues (structure mis-extracted?)";
        return false;
    }
    // b200sqp_evaluate perturbs the parameters like a Jacobian evaluation would; restore the caller's point
    return true;
}

SolverStatus SolverB200Lm::solve(OptimizationProblemInterface& problem, bool new_structure, bool new_run, double* obj_value)
{
    if (obj_value) *obj_value = -1;  // levenberg_marquardt_sparse.cpp:46
    if (!problem.isLeastSquaresProblem())
        return fail("cannot handle non-least-squares objectives or LS objectives in non-LS form.");  // :50-54
    std::vector<OptimizationProblemInterface*> one{&problem};
    std::vector<SolverStatus> st;
    std::vector<double> obj;
    (void)new_structure;  // the descriptor comparison in upload() decides whether device state must be rebuilt
    if (!solveBatch(one, new_run, &st, &obj)) return SolverStatus::Error;
    if (obj_value) *obj_value = obj[0];
    return st[0];
}

bool SolverB200Lm::solveBatch(const std::vector<OptimizationProblemInterface*>& problems, bool new_run, std::vector<SolverStatus>* statuses,
                              std::vector<double>* obj_values)
{
    const int B = (int)problems.size();
    if (B == 0) return true;
    // resetWeights / adaptWeights (levenberg_marquardt_sparse.cpp:83-86, 264-287) once per call: every object of the batch is in the
    // same OCP iteration.  The device gets the resulting weights as the initial weights of a new run.
    if (new_run || !_weights_initialised)
    {
        _w_eq = _opts.weight_eq, _w_ineq = _opts.weight_ineq, _w_bounds = _opts.weight_bounds;
    }
    else
    {
        _w_eq     = std::min(_w_eq * _opts.adapt_factor_eq, _opts.adapt_max_eq);
        _w_ineq   = std::min(_w_ineq * _opts.adapt_factor_ineq, _opts.adapt_max_ineq);
        _w_bounds = std::min(_w_bounds * _opts.adapt_factor_bounds, _opts.adapt_max_bounds);
    }
    _weights_initialised = true;
    b200sqp_lm_options opts = _opts;
    opts.weight_eq = _w_eq, opts.weight_ineq = _w_ineq, opts.weight_bounds = _w_bounds;

    // bucket by grid size (same OCP kind: the parameter dimension identifies N)
    std::map<int, std::vector<int>> groups;
    for (int i = 0; i < B; ++i) groups[problems[i]->getParameterDimension()].push_back(i);
    if (groups.size() == 1) return solveUniform(problems, opts, statuses, obj_values);
    if (statuses) statuses->assign(B, SolverStatus::Error);
    if (obj_values) obj_values->assign(B, -1.0);
    for (auto& g : groups)
    {
        std::shared_ptr<SolverB200Lm>& child = _by_size[g.first];
        if (!child) child = std::make_shared<SolverB200Lm>();
        child->_dynamics = _dynamics, child->_dynamics_parameters = _dynamics_parameters, child->_collocation = _collocation;
        child->_integrator = _integrator, child->_stage_cost = _stage_cost, child->_final_cost = _final_cost;
        child->_final_constraint = _final_constraint, child->_xref = _xref, child->_device = _device;
        std::vector<OptimizationProblemInterface*> sub;
        for (int i : g.second) sub.push_back(problems[i]);
        std::vector<SolverStatus> st;
        std::vector<double> obj;
        if (!child->solveUniform(sub, opts, &st, &obj))
        {
            fail("bucket of parameter dimension " + std::to_string(g.first) + ": " + child->lastError());
            return false;
        }
        for (size_t j = 0; j < g.second.size(); ++j)
        {
            if (statuses) (*statuses)[g.second[j]] = st[j];
            if (obj_values) (*obj_values)[g.second[j]] = obj[j];
        }
    }
    return true;
}

bool SolverB200Lm::solveUniform(const std::vector<OptimizationProblemInterface*>& problems, const b200sqp_lm_options& opts,
                                std::vector<SolverStatus>* statuses, std::vector<double>* obj_values)
{
    const int B = (int)problems.size();
    std::vector<double> x0_first, xref_first;
    if (!upload(*problems[0], B, &x0_first, &xref_first))
    {
        fail(_error);
        return false;
    }
    const int n = _dims.n_params, nx = _ocp.nx;
    // staging buffers live in the solver object: a fresh 10 MB of zeroed pages per call costs more than the device time of the batch
    std::vector<double>&x0 = _x0_buf, &xref = _xref_buf, &params = _params_buf;
    x0.resize((size_t)B * nx);
    xref.resize((size_t)B * nx);
    params.resize((size_t)B * n);
    std::copy(x0_first.begin(), x0_first.end(), x0.begin());
    std::copy(xref_first.begin(), xref_first.end(), xref.begin());
    std::atomic<int> first_bad(B);
    std::vector<std::string> errors(B > 0 ? 1 : 0);
    std::mutex error_mutex;
    forEachObject(B, [&](int i) {
        std::string err;
        if (i > 0 && !instanceData(*problems[i], x0.data() + (size_t)i * nx, xref.data() + (size_t)i * nx, err))
        {
            std::lock_guard<std::mutex> lock(error_mutex);
            if (i < first_bad.load())
            {
                first_bad = i;
                errors[0] = err;
            }
            return;
        }
        Eigen::Map<Eigen::VectorXd> p(params.data() + (size_t)i * n, n);
        problems[i]->getParameterVector(p);
    });
    if (first_bad.load() < B)
    {
        fail("problems of one batch must share one structure (object " + std::to_string(first_bad.load()) + "): " + errors[0]);
        return false;
    }
    if (b200sqp_set_problem_data(_handle, x0.data(), xref.data()) != 0 || !uploadReferenceTrajectory(B) ||
        b200sqp_set_params(_handle, params.data()) != 0)
    {
        fail(std::string("upload failed: ") + b200sqp_last_error());
        return false;
    }
    if (_fresh)
    {
        // new structure on the device: run the guard once, then restore the unperturbed parameters
        if (!selfCheck(*problems[0]) || b200sqp_set_params(_handle, params.data()) != 0)
        {
            fail(_error);
            return false;
        }
    }
    std::vector<int32_t>& status = _status_buf;
    std::vector<double>& chi2 = _chi2_buf;
    status.resize(B);
    chi2.resize(B);
    if (b200sqp_solve(_handle, &opts, 1, status.data(), chi2.data()) != 0 || b200sqp_get_params(_handle, params.data()) != 0)
    {
        fail(std::string("solve failed: ") + b200sqp_last_error());
        return false;
    }
    forEachObject(B, [&](int i) { problems[i]->setParameterVector(Eigen::Map<const Eigen::VectorXd>(params.data() + (size_t)i * n, n)); });
    if (statuses)
    {
        statuses->resize(B);
        for (int i = 0; i < B; ++i) (*statuses)[i] = status[i] == B200SQP_STATUS_CONVERGED ? SolverStatus::Converged : SolverStatus::EarlyTerminated;
    }
    if (obj_values) obj_values->assign(chi2.begin(), chi2.end());
    return true;
}

bool SolverB200Lm::evaluateOnDevice(OptimizationProblemInterface& problem, double weight_eq, double weight_ineq, double weight_bounds,
                                    Eigen::VectorXd* values, Eigen::SparseMatrix<double>* jacobian)
{
    if (!upload(problem, 1))
    {
        fail(_error);
        return false;
    }
    const int n = _dims.n_params, m = _dims.m_lsq + _dims.m_eq + _dims.m_ineq + _dims.m_bounds, nnz = _dims.nnz_jacobian;
    if (_fresh || (int)_col_ptr.size() != n + 1)
    {
        _col_ptr.assign(n + 1, 0);
        _row_idx.assign(nnz, 0);
        if (b200sqp_jacobian_pattern(&_ocp, _col_ptr.data(), _row_idx.data()) != 0)
        {
            fail(std::string("b200sqp_jacobian_pattern: ") + b200sqp_last_error());
            return false;
        }
    }
    std::vector<double> x0, xref;
    b200sqp_ocp d;
    if (!describe(problem, d, x0, xref))
    {
        fail(_error);
        return false;
    }
    Eigen::VectorXd params(n);
    problem.getParameterVector(params);
    if (b200sqp_set_problem_data(_handle, x0.data(), xref.data()) != 0 || !uploadReferenceTrajectory(1) || b200sqp_set_params(_handle, params.data()) != 0)
    {
        fail(std::string("upload failed: ") + b200sqp_last_error());
        return false;
    }
    if (_fresh)
    {
        // new structure on the device: run the guard once, then restore the unperturbed parameters
        if (!selfCheck(problem) || b200sqp_set_params(_handle, params.data()) != 0)
        {
            fail(_error);
            return false;
        }
    }
    std::vector<double> vals(m), jac(jacobian ? nnz : 0);
    if (b200sqp_evaluate(_handle, weight_eq, weight_ineq, weight_bounds, vals.data(), jacobian ? jac.data() : nullptr) != 0 ||
        b200sqp_get_params(_handle, params.data()) != 0)
    {
        fail(std::string("b200sqp_evaluate: ") + b200sqp_last_error());
        return false;
    }
    if (jacobian) problem.setParameterVector(params);  // the in-place differences leave their drift in the vertices, as in the reference
    if (values) *values = Eigen::Map<const Eigen::VectorXd>(vals.data(), m);
    if (jacobian)
    {
        // compressed CSC with the reference's pattern (explicit zeros included, rows ascending within a column)
        Eigen::Map<const Eigen::SparseMatrix<double, Eigen::ColMajor, int32_t>> view(m, n, nnz, _col_ptr.data(), _row_idx.data(), jac.data());
        *jacobian = view;
    }
    return true;
}

void HyperGraphOptimizationProblemB200::computeCombinedSparseJacobian(Eigen::SparseMatrix<double>& jacobian, bool objective_lsq, bool equality,
                                                                      bool inequality, bool finite_combined_bounds, bool active_ineq,
                                                                      double weight_eq, double weight_ineq, double weight_bounds,
                                                                      const Eigen::VectorXd* /*values*/, const Eigen::VectorXi* /*col_nnz*/)
{
    // the device evaluates the combined Jacobian the least-squares solvers ask for: all four categories, active-set inequality rows
    // (it recomputes the activity from the residuals at the same point, which is what the caller's `values` hold)
    if (!_evaluator || !(objective_lsq && equality && inequality && finite_combined_bounds && active_ineq))
    {
        PRINT_ERROR("HyperGraphOptimizationProblemB200: no device evaluator set, or a category selection the device path does not serve");
        _failed = true;
        jacobian.setZero();
        return;
    }
    if (!_evaluator->evaluateOnDevice(*this, weight_eq, weight_ineq, weight_bounds, nullptr, &jacobian))
    {
        PRINT_ERROR("HyperGraphOptimizationProblemB200: " << _evaluator->lastError());
        _failed = true;
        jacobian.setZero();
        return;
    }
    ++_device_evaluations;
}

}  // namespace corbo
