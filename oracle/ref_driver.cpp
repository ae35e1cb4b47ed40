// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// Driver around the UNMODIFIED reference (control_box_rst, compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libcorbo_ref.so).  It builds the OCP a b200sqp_ocp descriptor names out of the reference's own classes
// (StructuredOptimalControlProblem + FiniteDifferencesGrid/... + HyperGraphOptimizationProblemEdgeBased +
// LevenbergMarquardtSparse), runs it, and hands results back through a small C API used by
//   * tests/  (pinning oracle/sqp_oracle.cpp and generating tests/golden/ fixtures: tests/golden/make_golden.py),
//   * bench.py --impl reference / cpu_baseline (the reference's own CPU path timed on the host cores).
// Nothing here is linked into libb200sqp.so.
#include "../include/b200sqp.h"
#include "ref_models.h"

#include <corbo-core/reference_trajectory.h>
#include <corbo-core/time.h>
#include <corbo-numerics/explicit_integrators.h>
#include <corbo-numerics/finite_differences.h>
#include <corbo-numerics/finite_differences_collocation.h>
#include <corbo-optimal-control/functions/final_state_constraints.h>
#include <corbo-optimal-control/functions/final_state_cost.h>
#include <corbo-optimal-control/functions/minimum_time.h>
#include <corbo-optimal-control/functions/quadratic_cost.h>
#include <corbo-optimal-control/structured_ocp/discretization_grids/finite_differences_grid.h>
#include <corbo-optimal-control/structured_ocp/discretization_grids/multiple_shooting_grid.h>
#include <corbo-optimal-control/structured_ocp/discretization_grids/non_uniform_finite_differences_variable_grid.h>
#include <corbo-optimal-control/structured_ocp/structured_optimal_control_problem.h>
#include <corbo-optimization/hyper_graph/hyper_graph_optimization_problem_edge_based.h>
#include <corbo-optimization/simple_optimization_problem.h>
#include <corbo-optimization/solver/levenberg_marquardt_sparse.h>
#include <corbo-systems/benchmark/linear_benchmark_systems.h>
#include <corbo-systems/benchmark/nonlinear_benchmark_systems.h>
#include <corbo-systems/output_function_interface.h>
#include <corbo-plants/simulated_plant.h>
#include <corbo-controllers/predictive_controller.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

using namespace corbo;

namespace {

// Subclasses that only widen access to protected members (vertex containers); no behaviour is overridden.
struct FdGridProbe : public FiniteDifferencesGrid
{
    std::vector<VectorVertex>& xs() { return _x_seq; }
    std::vector<VectorVertex>& us() { return _u_seq; }
    PartiallyFixedVectorVertex& xf() { return _xf; }
    ScalarVertex& dt() { return _dt; }
};
struct NuGridProbe : public NonUniformFiniteDifferencesVariableGrid
{
    std::vector<VectorVertex>& xs() { return _x_seq; }
    std::vector<VectorVertex>& us() { return _u_seq; }
    std::vector<ScalarVertex>& dts() { return _dt_seq; }
    PartiallyFixedVectorVertex& xf() { return _xf; }
};
struct MsGridProbe : public MultipleShootingGrid
{
    std::vector<ShootingInterval>& intervals() { return _intervals; }
    PartiallyFixedVectorVertex& xf() { return _xf; }
    ScalarVertex& dt() { return _dt; }
};
struct OcpProbe : public StructuredOptimalControlProblem
{
    using StructuredOptimalControlProblem::StructuredOptimalControlProblem;
    BaseHyperGraphOptimizationProblem::Ptr problem() { return _optim_prob; }
    OptimizationEdgeSet::Ptr edges() { return _edges; }
    NlpSolverInterface::Ptr solver() { return _solver; }
    NlpFunctions& functions() { return _functions; }
};

// The hypergraph problem with an event log bolted on: every call the solver makes is forwarded unchanged to
// HyperGraphOptimizationProblemEdgeBased and recorded, which yields an iterate-level trace of the unmodified solver.
struct TracingProblem : public HyperGraphOptimizationProblemEdgeBased
{
    enum { EV_JACOBIAN = 0, EV_INCREMENT = 1, EV_RESTORE = 2, EV_DISCARD = 3 };
    struct Event
    {
        int type;
        double chi2;
        Eigen::VectorXd vec;
    };
    std::vector<Event> events;
    bool tracing = false;

    void computeCombinedSparseJacobian(Eigen::SparseMatrix<double>& jacobian, bool objective_lsq, bool equality, bool inequality,
                                       bool finite_combined_bounds, bool active_ineq, double weight_eq, double weight_ineq, double weight_bounds,
                                       const Eigen::VectorXd* values, const Eigen::VectorXi* col_nnz) override
    {
        if (tracing)
        {
            Event e;
            e.type = EV_JACOBIAN;
            e.chi2 = values ? values->squaredNorm() : -1;
            e.vec.resize(getParameterDimension());
            getParameterVector(e.vec);
            events.push_back(e);
        }
        HyperGraphOptimizationProblemEdgeBased::computeCombinedSparseJacobian(jacobian, objective_lsq, equality, inequality, finite_combined_bounds,
                                                                              active_ineq, weight_eq, weight_ineq, weight_bounds, values, col_nnz);
    }
    void applyIncrement(const Eigen::Ref<const Eigen::VectorXd>& increment) override
    {
        if (tracing) events.push_back({EV_INCREMENT, 0.0, increment});
        HyperGraphOptimizationProblemEdgeBased::applyIncrement(increment);
    }
    void restoreBackupParameters(bool keep_backup) override
    {
        if (tracing) events.push_back({EV_RESTORE, 0.0, Eigen::VectorXd()});
        HyperGraphOptimizationProblemEdgeBased::restoreBackupParameters(keep_backup);
    }
    void discardBackupParameters(bool all = false) override
    {
        if (tracing) events.push_back({EV_DISCARD, 0.0, Eigen::VectorXd()});
        HyperGraphOptimizationProblemEdgeBased::discardBackupParameters(all);
    }
};

struct RefOcp
{
    std::shared_ptr<OcpProbe> ocp;
    DiscretizationGridInterface::Ptr grid;
    std::shared_ptr<TracingProblem> problem;
    std::shared_ptr<LevenbergMarquardtSparse> solver;
    SystemDynamicsInterface::Ptr dynamics;
    ReferenceTrajectoryInterface::Ptr xref;  // StaticReference, or TableReference when a trajectory is given (corbo_ref_set_xref_points)
    ZeroReference::Ptr uref;
    int grid_kind = 0;
};

SystemDynamicsInterface::Ptr makeDynamics(const b200sqp_ocp& d)
{
    switch (d.dynamics)
    {
        case B200SQP_DYN_VAN_DER_POL:
        {
            auto s = std::make_shared<VanDerPolOscillator>();
            s->setDampingCoefficient(d.dyn_params[0]);
            return s;
        }
        case B200SQP_DYN_DUFFING:
        {
            auto s = std::make_shared<DuffingOscillator>();
            s->setParameters(d.dyn_params[0], d.dyn_params[1], d.dyn_params[2]);
            return s;
        }
        case B200SQP_DYN_SIMPLE_PENDULUM:
        {
            auto s = std::make_shared<SimplePendulum>();
            s->setParameters(d.dyn_params[0], d.dyn_params[1], d.dyn_params[2], d.dyn_params[3]);
            return s;
        }
        case B200SQP_DYN_CART_POLE:
            return std::make_shared<CartPole>();  // parameters are private constants in the reference
        case B200SQP_DYN_DOUBLE_INTEGRATOR:
        {
            auto s = std::make_shared<SerialIntegratorSystem>(2);
            s->setTimeConstant(d.dyn_params[0]);
            return s;
        }
        case B200SQP_DYN_FREE_SPACE_ROCKET:
            return std::make_shared<FreeSpaceRocket>();
        case B200SQP_DYN_MASSLESS_PENDULUM:
        {
            auto s = std::make_shared<MasslessPendulum>();
            s->setParameter(d.dyn_params[0]);
            return s;
        }
        case B200SQP_DYN_TOY_EXAMPLE:
        {
            auto s = std::make_shared<ToyExample>();
            s->setParameters(d.dyn_params[0]);
            return s;
        }
        case B200SQP_DYN_ARTSTEINS_CIRCLE:
            return std::make_shared<ArtsteinsCircle>();
        case B200SQP_DYN_LINEAR_2X1:
        case B200SQP_DYN_LINEAR_3X1:
        case B200SQP_DYN_LINEAR_4X1:
        case B200SQP_DYN_LINEAR_4X2:
        {
            auto s = std::make_shared<LinearStateSpaceModel>();
            Eigen::MatrixXd A = Eigen::Map<const Eigen::MatrixXd>(d.dyn_params, d.nx, d.nx);  // column-major
            Eigen::MatrixXd B = Eigen::Map<const Eigen::MatrixXd>(d.dyn_params + d.nx * d.nx, d.nx, d.nu);
            s->setParameters(A, B);
            return s;
        }
        case B200SQP_DYN_TRIPLE_INTEGRATOR:
        case B200SQP_DYN_QUAD_INTEGRATOR:
        {
            auto s = std::make_shared<SerialIntegratorSystem>(d.nx);
            s->setTimeConstant(d.dyn_params[0]);
            return s;
        }
        case B200SQP_DYN_UNICYCLE:
            return std::make_shared<b200ref::Unicycle>();
        case B200SQP_DYN_QUADROTOR:
            return std::make_shared<b200ref::Quadrotor>(d.dyn_params[0], d.dyn_params[1], d.dyn_params[2], d.dyn_params[3], d.dyn_params[4]);
    }
    return {};
}

FiniteDifferencesCollocationInterface::Ptr makeCollocation(int id)
{
    switch (id)
    {
        case B200SQP_COLL_FORWARD:
            return std::make_shared<ForwardDiffCollocation>();
        case B200SQP_COLL_BACKWARD:
            return std::make_shared<BackwardDiffCollocation>();
        case B200SQP_COLL_MIDPOINT:
            return std::make_shared<MidpointDiffCollocation>();
        default:
            return std::make_shared<CrankNicolsonDiffCollocation>();
    }
}

bool buildOcp(const b200sqp_ocp& d, const b200sqp_lm_options& o, RefOcp& r)
{
    r.dynamics = makeDynamics(d);
    if (!r.dynamics || r.dynamics->getStateDimension() != d.nx || r.dynamics->getInputDimension() != d.nu) return false;
    r.grid_kind = d.grid;

    Eigen::Matrix<bool, -1, 1> xf_fixed(d.nx);
    for (int i = 0; i < d.nx; ++i) xf_fixed[i] = d.xf_fixed[i] != 0;

    if (d.grid == B200SQP_GRID_FD_UNIFORM)
    {
        auto g = std::make_shared<FdGridProbe>();
        g->setNRef(d.n_grid);
        g->setDtRef(d.dt_ref);
        g->setFiniteDifferencesCollocationMethod(makeCollocation(d.collocation));
        g->setCostIntegrationRule(FullDiscretizationGridBase::CostIntegrationRule::LeftSum);  // member is otherwise uninitialised
        g->setXfFixed(xf_fixed);
        r.grid = g;
    }
    else if (d.grid == B200SQP_GRID_FD_NONUNIFORM_VARDT)
    {
        auto g = std::make_shared<NuGridProbe>();
        g->setNRef(d.n_grid);
        g->setDtRef(d.dt_ref);
        g->setDtBounds(d.dt_lb, d.dt_ub);
        g->disableGridAdaptation();
        g->setDtEqConstraint(d.dt_eq_constraint != 0);
        g->setFiniteDifferencesCollocationMethod(makeCollocation(d.collocation));
        g->setCostIntegrationRule(NonUniformFullDiscretizationGridBase::CostIntegrationRule::LeftSum);
        g->setXfFixed(xf_fixed);
        r.grid = g;
    }
    else if (d.grid == B200SQP_GRID_MULTIPLE_SHOOTING)
    {
        auto g = std::make_shared<MsGridProbe>();
        g->setNRef(d.n_grid);
        g->setDtRef(d.dt_ref);
        g->setNumControlsPerShootingInterval(1);
        if (d.integrator == B200SQP_INT_RK4)
            g->setNumericalIntegrator(std::make_shared<IntegratorExplicitRungeKutta4>());
        else
            g->setNumericalIntegrator(std::make_shared<IntegratorExplicitEuler>());
        g->setXfFixed(xf_fixed);
        r.grid = g;
    }
    else
        return false;

    r.problem = std::make_shared<TracingProblem>();
    r.solver  = std::make_shared<LevenbergMarquardtSparse>();
    r.solver->setIterations(o.iterations);
    r.solver->setPenaltyWeights(o.weight_eq, o.weight_ineq, o.weight_bounds);
    r.solver->setWeightAdapation(o.adapt_factor_eq, o.adapt_factor_ineq, o.adapt_factor_bounds, o.adapt_max_eq, o.adapt_max_ineq,
                                 o.adapt_max_bounds);

    r.ocp = std::make_shared<OcpProbe>(r.grid, r.dynamics, r.problem, r.solver);

    if (d.stage_cost == B200SQP_COST_QUADRATIC_LSQ)
    {
        Eigen::MatrixXd Q = Eigen::MatrixXd::Zero(d.nx, d.nx), R = Eigen::MatrixXd::Zero(d.nu, d.nu);
        for (int i = 0; i < d.nx; ++i) Q(i, i) = d.q_diag[i];
        for (int i = 0; i < d.nu; ++i) R(i, i) = d.r_diag[i];
        if (d.q_dense) Q = Eigen::Map<const Eigen::Matrix<double, -1, -1, Eigen::RowMajor>>(d.q_full, d.nx, d.nx);
        if (d.r_dense) R = Eigen::Map<const Eigen::Matrix<double, -1, -1, Eigen::RowMajor>>(d.r_full, d.nu, d.nu);
        r.ocp->setStageCost(std::make_shared<QuadraticFormCost>(Q, R, false, true));
    }
    else if (d.stage_cost == B200SQP_COST_MINIMUM_TIME_LSQ)
    {
        r.ocp->setStageCost(std::make_shared<MinimumTime>(true));
    }
    if (d.final_cost == 1)
    {
        Eigen::MatrixXd Qf = Eigen::MatrixXd::Zero(d.nx, d.nx);
        for (int i = 0; i < d.nx; ++i) Qf(i, i) = d.qf_diag[i];
        if (d.qf_dense) Qf = Eigen::Map<const Eigen::Matrix<double, -1, -1, Eigen::RowMajor>>(d.qf_full, d.nx, d.nx);
        r.ocp->setFinalStageCost(std::make_shared<QuadraticFinalStateCost>(Qf, true));
    }

    if (d.final_constraint == B200SQP_FINAL_CONSTRAINT_EQUALITY)
    {
        Eigen::VectorXd tx(d.nx);
        for (int i = 0; i < d.nx; ++i) tx[i] = d.term_xref[i];
        r.ocp->setFinalStageConstraint(std::make_shared<TerminalEqualityConstraint>(tx));
    }
    else if (d.final_constraint == B200SQP_FINAL_CONSTRAINT_BALL)
    {
        Eigen::MatrixXd S = Eigen::MatrixXd::Zero(d.nx, d.nx);
        for (int i = 0; i < d.nx; ++i) S(i, i) = d.term_s_diag[i];
        r.ocp->setFinalStageConstraint(std::make_shared<TerminalBall>(S, d.term_gamma));
    }

    Eigen::VectorXd xlb(d.nx), xub(d.nx), ulb(d.nu), uub(d.nu);
    for (int i = 0; i < d.nx; ++i)
    {
        xlb[i] = d.x_lb[i];
        xub[i] = d.x_ub[i];
    }
    for (int i = 0; i < d.nu; ++i)
    {
        ulb[i] = d.u_lb[i];
        uub[i] = d.u_ub[i];
    }
    r.ocp->setBounds(xlb, xub, ulb, uub);

    r.uref = std::make_shared<ZeroReference>(d.nu);
    return r.ocp->initialize();
}

// Everything StructuredOptimalControlProblem::compute (structured_optimal_control_problem.cpp:77-154) does before
// _solver->solve(): grid update (initialises the trajectories on the first call) and index precomputation.
// A user-side reference trajectory for the reference's ReferenceTrajectoryInterface (core/reference_trajectory.h:60-95): non-static, its
// cached value at grid point k is row k of a table.  The reference's own non-static classes (DiscreteTimeReferenceTrajectory, ...)
// produce such rows by interpolating a time series; handing the rows over directly keeps the checkers' inputs bit-identical.
class TableReference : public ReferenceTrajectoryInterface
{
 public:
    TableReference(const double* rows, int n, int dim) : _rows(n, OutputVector(dim))
    {
        for (int k = 0; k < n; ++k)
            for (int i = 0; i < dim; ++i) _rows[k][i] = rows[(size_t)k * dim + i];
    }
    Ptr getInstance() const override { return std::make_shared<TableReference>(nullptr, 0, 0); }
    bool isStatic() const override { return false; }
    bool isZero() const override { return false; }
    int getDimension() const override { return _rows.empty() ? 0 : (int)_rows[0].size(); }
    void precompute(double, int, Time) override {}
    void precompute(const std::vector<double>&, Time) override {}
    void getReference(const Time&, OutputVector& ref) const override { ref = _rows.front(); }
    const OutputVector& getReferenceCached(int k) const override { return _rows[std::min<int>(std::max(k, 0), (int)_rows.size() - 1)]; }
    const OutputVector& getNextSteadyState(const Time&) override { return _rows.back(); }
    bool isCached(double, int, Time) const override { return true; }
    bool isCached(const std::vector<double>&, Time) const override { return true; }

 private:
    std::vector<OutputVector> _rows;
};

// see sqp_oracle_set_xref_points: > 1 = every `xref` argument is a trajectory [n_grid][nx] (row k = getReferenceCached(k))
static int g_xref_points = 0;
static int xrefStride(const b200sqp_ocp& d) { return g_xref_points > 1 ? d.n_grid * d.nx : d.nx; }

bool prepare(RefOcp& r, const b200sqp_ocp& d, const double* x0, const double* xref, bool new_run, bool* structure_changed)
{
    Eigen::VectorXd x0v = Eigen::Map<const Eigen::VectorXd>(x0, d.nx);
    if (g_xref_points > 1 && xref)
    {
        if (g_xref_points != d.n_grid) return false;
        r.xref = std::make_shared<TableReference>(xref, d.n_grid, d.nx);
    }
    else
    {
        Eigen::VectorXd xr = xref ? Eigen::VectorXd(Eigen::Map<const Eigen::VectorXd>(xref, d.nx)) : Eigen::VectorXd(Eigen::VectorXd::Zero(d.nx));
        r.xref             = std::make_shared<StaticReference>(xr);
    }
    Eigen::VectorXd uprev = Eigen::VectorXd::Zero(d.nu);
    GridUpdateResult res  = r.grid->update(x0v, *r.xref, *r.uref, r.ocp->functions(), *r.ocp->edges(), r.dynamics, new_run, Time(0), nullptr, &uprev,
                                           r.grid->getInitialDt(), nullptr, nullptr);
    if (res.vertices_updated) r.problem->precomputeVertexQuantities();
    if (res.updated()) r.problem->precomputeEdgeQuantities();
    if (structure_changed) *structure_changed = res.updated();
    return true;
}

void fillDims(RefOcp& r, b200sqp_dims* out)
{
    OptimizationProblemInterface& p = *r.problem;
    std::memset(out, 0, sizeof(*out));
    out->n_params     = p.getParameterDimension();
    out->m_lsq        = p.getLsqObjectiveDimension();
    out->m_eq         = p.getEqualityDimension();
    out->m_ineq       = p.getInequalityDimension();
    out->m_bounds     = p.finiteCombinedBoundsDimension();
    out->nnz_jacobian = p.computeSparseJacobianLsqObjectiveNNZ() + p.computeSparseJacobianEqualitiesNNZ() + p.computeSparseJacobianInequalitiesNNZ() +
                        p.computeSparseJacobianFiniteCombinedBoundsNNZ();
    // structural nnz of triu(J^T J): evaluate J once (a copy of the parameters is restored afterwards)
    const int m = out->m_lsq + out->m_eq + out->m_ineq + out->m_bounds, n = out->n_params;
    Eigen::VectorXd backup(n);
    p.getParameterVector(backup);
    Eigen::SparseMatrix<double> J(m, n);
    p.computeCombinedSparseJacobian(J, true, true, true, true, false, 1.0, 1.0, 1.0, nullptr, nullptr);
    p.setParameterVector(backup);
    Eigen::SparseMatrix<double> H = J.transpose() * J;
    int nnz_full                  = H.nonZeros();
    out->nnz_hessian_upper        = (nnz_full + n) / 2;
    const int64_t s               = 8;
    out->algorithmic_bytes_per_iteration =
        s * (2 * ((int64_t)J.nonZeros() + 2 * (int64_t)out->nnz_hessian_upper + 2 * (int64_t)m + 2 * (int64_t)n) + 4 * (int64_t)n);
}

int vertexIdxOrMinus1(const VertexInterface& v) { return v.getDimensionUnfixed() > 0 ? v.getVertexIdx() : -1; }

}  // namespace

extern "C" {

int corbo_ref_dims(const b200sqp_ocp* d, b200sqp_dims* out)
{
    b200sqp_lm_options o = {10, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    std::vector<double> x0(d->nx, 0.25);
    if (!prepare(r, *d, x0.data(), nullptr, true, nullptr)) return -1;
    fillDims(r, out);
    return 0;
}

// x_idx [N], u_idx [N-1], dt_idx [N-1]: parameter index of the vertex' first free component, -1 if the vertex is fixed
int corbo_ref_vertex_indices(const b200sqp_ocp* d, int32_t* x_idx, int32_t* u_idx, int32_t* dt_idx)
{
    b200sqp_lm_options o = {10, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    std::vector<double> x0(d->nx, 0.25);
    if (!prepare(r, *d, x0.data(), nullptr, true, nullptr)) return -1;
    const int N = d->n_grid;
    if (d->grid == B200SQP_GRID_FD_UNIFORM)
    {
        auto* g = static_cast<FdGridProbe*>(r.grid.get());
        for (int k = 0; k < N - 1; ++k)
        {
            x_idx[k]  = vertexIdxOrMinus1(g->xs()[k]);
            u_idx[k]  = vertexIdxOrMinus1(g->us()[k]);
            dt_idx[k] = vertexIdxOrMinus1(g->dt());
        }
        x_idx[N - 1] = vertexIdxOrMinus1(g->xf());
    }
    else if (d->grid == B200SQP_GRID_FD_NONUNIFORM_VARDT)
    {
        auto* g = static_cast<NuGridProbe*>(r.grid.get());
        for (int k = 0; k < N - 1; ++k)
        {
            x_idx[k]  = vertexIdxOrMinus1(g->xs()[k]);
            u_idx[k]  = vertexIdxOrMinus1(g->us()[k]);
            dt_idx[k] = vertexIdxOrMinus1(g->dts()[k]);
        }
        x_idx[N - 1] = vertexIdxOrMinus1(g->xf());
    }
    else
    {
        auto* g = static_cast<MsGridProbe*>(r.grid.get());
        for (int k = 0; k < N - 1; ++k)
        {
            x_idx[k]  = vertexIdxOrMinus1(g->intervals()[k].s);
            u_idx[k]  = vertexIdxOrMinus1(g->intervals()[k].u_seq[0]);
            dt_idx[k] = vertexIdxOrMinus1(g->dt());
        }
        x_idx[N - 1] = vertexIdxOrMinus1(g->xf());
    }
    return 0;
}

// Edge row offsets per category in creation order: for every lsq / equality edge (dimension, edge index, attached vertex indices)
// flattened as: dim, idx, n_vertices, vertex_idx[0..3] (-1 padded) -> 7 ints per edge.  Returns the number of edges written.
int corbo_ref_edge_table(const b200sqp_ocp* d, int category /*0 lsq, 1 eq, 2 ineq*/, int32_t* table, int max_edges)
{
    b200sqp_lm_options o = {10, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    std::vector<double> x0(d->nx, 0.25);
    if (!prepare(r, *d, x0.data(), nullptr, true, nullptr)) return -1;
    OptimizationEdgeSet::Ptr es = r.ocp->edges();
    std::vector<BaseEdge::Ptr>* list =
        category == 0 ? &es->getLsqObjectiveEdgesRef() : (category == 1 ? &es->getEqualityEdgesRef() : &es->getInequalityEdgesRef());
    int cnt = 0;
    for (BaseEdge::Ptr& e : *list)
    {
        if (cnt >= max_edges) break;
        int32_t* row = table + 7 * cnt;
        row[0]       = e->getDimension();
        row[1]       = e->getEdgeIdx();
        row[2]       = e->getNumVertices();
        for (int v = 0; v < 4; ++v) row[3 + v] = v < e->getNumVertices() ? vertexIdxOrMinus1(*e->getVertexRaw(v)) : -1;
        ++cnt;
    }
    return cnt;
}

// Reference initial guess (FullDiscretizationGridBase::initializeSequences) as a parameter vector [n]
int corbo_ref_set_xref_points(int n_points)
{
    g_xref_points = n_points;
    return 0;
}

int corbo_ref_initial_params(const b200sqp_ocp* d, const double* x0, const double* xref, double* params)
{
    b200sqp_lm_options o = {10, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    if (!prepare(r, *d, x0, xref, true, nullptr)) return -1;
    Eigen::VectorXd p(r.problem->getParameterDimension());
    r.problem->getParameterVector(p);
    std::memcpy(params, p.data(), sizeof(double) * p.size());
    return 0;
}

// LevenbergMarquardtSparse::computeValues + computeCombinedSparseJacobian at `params` (NULL = reference initial guess).
// values [m]; jac_dense [m*n] row-major (zeros where structurally empty); jac_pattern [m*n] 1 where the sparse matrix stores an
// entry (explicit zeros included); params_after [n] = parameters after the in-place FD perturbations.
int corbo_ref_evaluate(const b200sqp_ocp* d, const double* x0, const double* xref, const double* params, double w_eq, double w_ineq, double w_b,
                       double* values, double* jac_dense, uint8_t* jac_pattern, double* params_after)
{
    b200sqp_lm_options o = {10, w_eq, w_ineq, w_b, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    if (!prepare(r, *d, x0, xref, true, nullptr)) return -1;
    OptimizationProblemInterface& p = *r.problem;
    const int n                     = p.getParameterDimension();
    if (params) p.setParameterVector(Eigen::Map<const Eigen::VectorXd>(params, n));
    const int m_lsq = p.getLsqObjectiveDimension(), m_eq = p.getEqualityDimension(), m_ineq = p.getInequalityDimension(),
              m_b = p.finiteCombinedBoundsDimension();
    const int m   = m_lsq + m_eq + m_ineq + m_b;
    Eigen::VectorXd v(m);
    // LevenbergMarquardtSparse::computeValues is protected; these are its five statements (levenberg_marquardt_sparse.cpp:222-246)
    int idx = 0;
    if (m_lsq > 0)
    {
        p.computeValuesLsqObjective(v.segment(idx, m_lsq));
        idx += m_lsq;
    }
    if (m_eq > 0)
    {
        p.computeValuesEquality(v.segment(idx, m_eq));
        v.segment(idx, m_eq) *= w_eq;
        idx += m_eq;
    }
    if (m_ineq > 0)
    {
        p.computeValuesActiveInequality(v.segment(idx, m_ineq), w_ineq);
        idx += m_ineq;
    }
    if (m_b > 0)
    {
        p.computeDistanceFiniteCombinedBounds(v.segment(idx, m_b));
        v.segment(idx, m_b) *= w_b;
    }
    if (values) std::memcpy(values, v.data(), sizeof(double) * m);
    if (jac_dense || jac_pattern || params_after)
    {
        Eigen::SparseMatrix<double> J(m, n);
        p.computeCombinedSparseJacobian(J, true, true, true, true, true, w_eq, w_ineq, w_b, &v, nullptr);
        if (jac_dense) std::memset(jac_dense, 0, sizeof(double) * m * n);
        if (jac_pattern) std::memset(jac_pattern, 0, (size_t)m * n);
        for (int c = 0; c < J.outerSize(); ++c)
            for (Eigen::SparseMatrix<double>::InnerIterator it(J, c); it; ++it)
            {
                if (jac_dense) jac_dense[(size_t)it.row() * n + it.col()] = it.value();
                if (jac_pattern) jac_pattern[(size_t)it.row() * n + it.col()] = 1;
            }
        if (params_after)
        {
            Eigen::VectorXd pa(n);
            p.getParameterVector(pa);
            std::memcpy(params_after, pa.data(), sizeof(double) * n);
        }
    }
    return 0;
}

// One instance through the unmodified solver with the event log on.
//  ev_type [max_events], ev_chi2 [max_events], ev_vec [max_events*n] (increment for EV_INCREMENT, parameters for EV_JACOBIAN)
int corbo_ref_trace(const b200sqp_ocp* d, const b200sqp_lm_options* o, const double* x0, const double* xref, const double* params_in,
                    double* params_out, double* chi2_out, int32_t* status_out, int max_events, int32_t* ev_type, double* ev_chi2, double* ev_vec,
                    int32_t* n_events)
{
    RefOcp r;
    if (!buildOcp(*d, *o, r)) return -1;
    bool changed = false;
    if (!prepare(r, *d, x0, xref, true, &changed)) return -1;
    const int n = r.problem->getParameterDimension();
    if (params_in) r.problem->setParameterVector(Eigen::Map<const Eigen::VectorXd>(params_in, n));
    r.problem->tracing = true;
    double obj         = -1;
    SolverStatus st    = r.solver->solve(*r.problem, true, true, &obj);
    r.problem->tracing = false;
    Eigen::VectorXd p(n);
    r.problem->getParameterVector(p);
    if (params_out) std::memcpy(params_out, p.data(), sizeof(double) * n);
    if (chi2_out) *chi2_out = obj;
    if (status_out) *status_out = (int)st;
    int cnt = 0;
    for (auto& e : r.problem->events)
    {
        if (cnt >= max_events) break;
        ev_type[cnt] = e.type;
        ev_chi2[cnt] = e.chi2;
        if (ev_vec)
        {
            std::memset(ev_vec + (size_t)cnt * n, 0, sizeof(double) * n);
            if (e.vec.size() == n) std::memcpy(ev_vec + (size_t)cnt * n, e.vec.data(), sizeof(double) * n);
        }
        ++cnt;
    }
    if (n_events) *n_events = (int)r.problem->events.size();
    return 0;
}

// A batch of independent instances through StructuredOptimalControlProblem::compute (cold: fresh objects per instance, the
// fairness rule of SURVEY.md section 8d), statically partitioned over `threads` std::threads (reference objects are
// single-thread-affine, nothing is shared).  x0 [B*nx], xref [B*nx] or NULL, params_in [B*n] or NULL (reference initial guess),
// params_out [B*n], chi2 [B], status [B] (any may be NULL).  Returns timings: seconds[0] = wall time of the whole batch,
// seconds[1] = sum over instances of the time spent inside solver->solve(), seconds[2] = sum of preparation times.
int corbo_ref_solve_batch(const b200sqp_ocp* d, const b200sqp_lm_options* o, int batch, const double* x0, const double* xref,
                          const double* params_in, double* params_out, double* chi2, int32_t* status, int threads, double* seconds)
{
    if (threads < 1) threads = 1;
    b200sqp_dims dims;
    if (corbo_ref_dims(d, &dims) != 0) return -1;
    const int n = dims.n_params;
    std::atomic<int> failures(0);
    std::vector<double> t_solve(threads, 0.0), t_prep(threads, 0.0);
    auto t_begin = std::chrono::steady_clock::now();
    auto worker  = [&](int tid) {
        const int lo = (int)((int64_t)batch * tid / threads), hi = (int)((int64_t)batch * (tid + 1) / threads);
        for (int i = lo; i < hi; ++i)
        {
            RefOcp r;
            if (!buildOcp(*d, *o, r))
            {
                ++failures;
                continue;
            }
            auto t0      = std::chrono::steady_clock::now();
            bool changed = false;
            prepare(r, *d, x0 + (size_t)i * d->nx, xref ? xref + (size_t)i * xrefStride(*d) : nullptr, true, &changed);
            if (params_in) r.problem->setParameterVector(Eigen::Map<const Eigen::VectorXd>(params_in + (size_t)i * n, n));
            auto t1         = std::chrono::steady_clock::now();
            double obj      = -1;
            SolverStatus st = r.solver->solve(*r.problem, changed, true, &obj);
            auto t2         = std::chrono::steady_clock::now();
            t_prep[tid] += std::chrono::duration<double>(t1 - t0).count();
            t_solve[tid] += std::chrono::duration<double>(t2 - t1).count();
            if (params_out)
            {
                Eigen::VectorXd p(n);
                r.problem->getParameterVector(p);
                std::memcpy(params_out + (size_t)i * n, p.data(), sizeof(double) * n);
            }
            if (chi2) chi2[i] = obj;
            if (status) status[i] = (int)st;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    auto t_end = std::chrono::steady_clock::now();
    if (seconds)
    {
        seconds[0] = std::chrono::duration<double>(t_end - t_begin).count();
        seconds[1] = 0;
        seconds[2] = 0;
        for (int t = 0; t < threads; ++t)
        {
            seconds[1] += t_solve[t];
            seconds[2] += t_prep[t];
        }
    }
    return failures.load() == 0 ? 0 : -1;
}

// Closed-loop plumbing check (BASELINE.json configs[0]): `steps` MPC steps of one instance, warm-started
// (StructuredOptimalControlProblem::compute with new_run = true per step, structure kept), plant = the same dynamics integrated
// with one explicit RK4 step of dt_ref.  u_applied [steps*nu], x_closed [ (steps+1)*nx ].
int corbo_ref_closed_loop(const b200sqp_ocp* d, const b200sqp_lm_options* o, const double* x0, int steps, double* u_applied, double* x_closed)
{
    RefOcp r;
    if (!buildOcp(*d, *o, r)) return -1;
    StaticReference xref(Eigen::VectorXd::Zero(d->nx));
    ZeroReference uref(d->nu);
    Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x0, d->nx);
    IntegratorExplicitRungeKutta4 rk4;
    std::memcpy(x_closed, x.data(), sizeof(double) * d->nx);
    for (int s = 0; s < steps; ++s)
    {
        if (!r.ocp->compute(x, xref, uref, nullptr, Time(s * d->dt_ref), true)) return -2;
        Eigen::VectorXd u(d->nu);
        if (!r.ocp->getFirstControlInput(u)) return -3;
        std::memcpy(u_applied + (size_t)s * d->nu, u.data(), sizeof(double) * d->nu);
        Eigen::VectorXd xn(d->nx);
        rk4.solveIVP(x, u, d->dt_ref, *r.dynamics, xn);
        x = xn;
        std::memcpy(x_closed + (size_t)(s + 1) * d->nx, x.data(), sizeof(double) * d->nx);
    }
    return 0;
}

// setWarmStart of whichever grid base the descriptor built (FullDiscretizationGridBase / ShootingGridBase / NonUniformFullDiscretizationGridBase)
static bool enableWarmStart(RefOcp& r)
{
    if (auto* g = dynamic_cast<FullDiscretizationGridBase*>(r.grid.get()))
    {
        g->setWarmStart(true);
        return true;
    }
    if (auto* g = dynamic_cast<ShootingGridBase*>(r.grid.get()))
    {
        g->setWarmStart(true);
        return true;
    }
    if (auto* g = dynamic_cast<NonUniformFullDiscretizationGridBase*>(r.grid.get()))
    {
        g->setWarmStart(true);
        return true;
    }
    return false;
}

// Moving-horizon warm start of the reference grid (FullDiscretizationGridBase::update -> warmStartShifting + findNearestState,
// full_discretization_grid_base.cpp:95-107,230-318), isolated: initialise at x0_old, overwrite the parameters with params_in, run the
// grid update for a new run at x0_new with warm start active, return the shifted parameters.
int corbo_ref_warm_start_shift(const b200sqp_ocp* d, const double* x0_old, const double* x0_new, const double* xref, const double* params_in,
                               double* params_out)
{
    b200sqp_lm_options o = {0, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    if (!enableWarmStart(r)) return -4;
    if (!prepare(r, *d, x0_old, xref, true, nullptr)) return -2;
    const int n = r.problem->getParameterDimension();
    r.problem->setParameterVector(Eigen::Map<const Eigen::VectorXd>(params_in, n));
    if (!prepare(r, *d, x0_new, xref, true, nullptr)) return -3;
    Eigen::VectorXd p(n);
    r.problem->getParameterVector(p);
    std::memcpy(params_out, p.data(), sizeof(double) * n);
    return 0;
}

// Closed loop like corbo_ref_closed_loop, with the grid's moving-horizon warm start switched on (setWarmStart(true))
int corbo_ref_closed_loop_shift(const b200sqp_ocp* d, const b200sqp_lm_options* o, const double* x0, int steps, double* u_applied, double* x_closed)
{
    RefOcp r;
    if (!buildOcp(*d, *o, r)) return -1;
    if (!enableWarmStart(r)) return -4;
    StaticReference xref(Eigen::VectorXd::Zero(d->nx));
    ZeroReference uref(d->nu);
    Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x0, d->nx);
    IntegratorExplicitRungeKutta4 rk4;
    std::memcpy(x_closed, x.data(), sizeof(double) * d->nx);
    for (int s = 0; s < steps; ++s)
    {
        if (!r.ocp->compute(x, xref, uref, nullptr, Time(s * d->dt_ref), true)) return -2;
        Eigen::VectorXd u(d->nu);
        if (!r.ocp->getFirstControlInput(u)) return -3;
        std::memcpy(u_applied + (size_t)s * d->nu, u.data(), sizeof(double) * d->nu);
        Eigen::VectorXd xn(d->nx);
        rk4.solveIVP(x, u, d->dt_ref, *r.dynamics, xn);
        x = xn;
        std::memcpy(x_closed + (size_t)(s + 1) * d->nx, x.data(), sizeof(double) * d->nx);
    }
    return 0;
}

// The reference's own plant: SimulatedPlant (FullStateSystemOutput, no disturbances / dead time) stepped once per point with
// PlantInterface::control(u, dt, t) -> SimulatedPlant::control(u_sequence, ...) -> _integrator->solveIVP.
// integrator 0 = IntegratorExplicitEuler (the plant's default), 1 = IntegratorExplicitRungeKutta4.
static std::shared_ptr<SimulatedPlant> makePlant(const b200sqp_ocp& d, int integrator)
{
    SystemDynamicsInterface::Ptr dyn = makeDynamics(d);
    if (!dyn) return {};
    auto plant = std::make_shared<SimulatedPlant>(dyn, std::make_shared<FullStateSystemOutput>());
    if (integrator == 1) plant->setIntegrator(std::make_shared<IntegratorExplicitRungeKutta4>());
    return plant;
}

int corbo_ref_plant_step(const b200sqp_ocp* d, int integrator, double dt, int batch, const double* x, const double* u, double* x_next)
{
    const int nx = d->nx, nu = d->nu;
    for (int i = 0; i < batch; ++i)
    {
        auto plant = makePlant(*d, integrator);  // fresh plant per point: the control buffer is stateful
        if (!plant) return -1;
        if (!plant->setState(Eigen::Map<const Eigen::VectorXd>(x + (size_t)i * nx, nx))) return -2;
        Eigen::VectorXd ui = Eigen::Map<const Eigen::VectorXd>(u + (size_t)i * nu, nu);
        if (!static_cast<PlantInterface&>(*plant).control(ui, Duration(dt), Time(0))) return -3;
        Eigen::VectorXd y(nx);
        if (!plant->output(y, Time(dt))) return -4;
        std::memcpy(x_next + (size_t)i * nx, y.data(), sizeof(double) * nx);
    }
    return 0;
}

// The closed loop of ClosedLoopControlTask::performTask (task_closed_loop_control.cpp:153-235) for one instance with the reference's own
// classes: plant->output -> PredictiveController::step (one OCP iteration) -> plant->control(u_sequence, x_sequence, dt, t).
// warm_start != 0 switches the grid's moving-horizon warm start on.  u_applied [steps*nu], x_closed [(steps+1)*nx].
int corbo_ref_closed_loop_plant(const b200sqp_ocp* d, const b200sqp_lm_options* o, const double* x0, int steps, int integrator, double plant_dt,
                                int warm_start, double* u_applied, double* x_closed)
{
    RefOcp r;
    if (!buildOcp(*d, *o, r)) return -1;
    if (warm_start)
    {
        if (!enableWarmStart(r)) return -4;
    }
    auto plant = makePlant(*d, integrator);
    if (!plant) return -5;
    if (!plant->setState(Eigen::Map<const Eigen::VectorXd>(x0, d->nx))) return -6;
    PredictiveController controller;
    controller.setOptimalControlProblem(r.ocp);
    controller.setNumOcpIterations(1);
    controller.setAutoUpdatePreviousControl(false);
    StaticReference xref(Eigen::VectorXd::Zero(d->nx));
    ZeroReference uref(d->nu);
    Eigen::VectorXd x(d->nx);
    const Duration dt(plant_dt);
    Time t(0);
    for (int s = 0; s < steps; ++s)
    {
        if (!plant->output(x, t)) return -7;
        std::memcpy(x_closed + (size_t)s * d->nx, x.data(), sizeof(double) * d->nx);
        TimeSeries::Ptr u_seq = std::make_shared<TimeSeries>(), x_seq = std::make_shared<TimeSeries>();
        if (!controller.step(x, xref, uref, dt, t, u_seq, x_seq)) return -2;
        if (u_seq->getTimeDimension() < 1) return -3;
        Eigen::VectorXd u = u_seq->getValuesMap(0);
        std::memcpy(u_applied + (size_t)s * d->nu, u.data(), sizeof(double) * d->nu);
        if (!plant->control(u_seq, x_seq, dt, t)) return -8;
        t += dt;
    }
    if (!plant->output(x, t)) return -7;
    std::memcpy(x_closed + (size_t)steps * d->nx, x.data(), sizeof(double) * d->nx);
    return 0;
}

// One call of the reference's grid adaptation in isolation: initialise the grid (N = n_grid), overwrite its vertices with the given
// trajectory (x_in [N][nx] incl. start and final state, u_in [N-1][nu], dt_in [N-1]), run the grid update of a continued run
// (new_run = false, warm start on -> adaptGrid, no re-initialisation) and read the vertices back (capacity N + 1 rows); *n_out = new N.
// strategy 0: setGridAdaptTimeBasedSingleStep(n_max, p0 = dt_hyst_ratio); 1: setGridAdaptRedundantControls(n_max, p1 = num_backup_nodes, p0 = epsilon)
static void selectGridAdaptation(NuGridProbe* g, int strategy, int n_min, int n_max, double p0, int p1)
{
    if (strategy == 1)
        g->setGridAdaptRedundantControls(n_max, p1, p0);
    else
        g->setGridAdaptTimeBasedSingleStep(n_max, p0);
    g->setNmin(n_min);
}

static int adaptOnce(const b200sqp_ocp* d, int strategy, int n_min, int n_max, double p0, int p1, const double* x_in, const double* u_in,
                     const double* dt_in, double* x_out, double* u_out, double* dt_out, int32_t* n_out)
{
    if (d->grid != B200SQP_GRID_FD_NONUNIFORM_VARDT) return -9;
    b200sqp_lm_options o = {0, 2, 2, 2, 1, 1, 1, 500, 500, 500};
    RefOcp r;
    if (!buildOcp(*d, o, r)) return -1;
    auto* g = dynamic_cast<NuGridProbe*>(r.grid.get());
    if (!g) return -9;
    selectGridAdaptation(g, strategy, n_min, n_max, p0, p1);
    g->setWarmStart(true);
    const int N = d->n_grid;
    if (!prepare(r, *d, x_in, x_in + (size_t)(N - 1) * d->nx, true, nullptr)) return -2;
    for (int k = 0; k < N - 1; ++k)
    {
        g->xs()[k].values() = Eigen::Map<const Eigen::VectorXd>(x_in + (size_t)k * d->nx, d->nx);
        g->us()[k].values() = Eigen::Map<const Eigen::VectorXd>(u_in + (size_t)k * d->nu, d->nu);
        g->dts()[k].value() = dt_in[k];
    }
    g->xf().values() = Eigen::Map<const Eigen::VectorXd>(x_in + (size_t)(N - 1) * d->nx, d->nx);
    if (!prepare(r, *d, x_in, x_in + (size_t)(N - 1) * d->nx, false, nullptr)) return -3;
    const int n = g->getN();
    *n_out      = n;
    for (int k = 0; k < n - 1; ++k)
    {
        std::memcpy(x_out + (size_t)k * d->nx, g->xs()[k].values().data(), sizeof(double) * d->nx);
        std::memcpy(u_out + (size_t)k * d->nu, g->us()[k].values().data(), sizeof(double) * d->nu);
        dt_out[k] = g->dts()[k].value();
    }
    std::memcpy(x_out + (size_t)(n - 1) * d->nx, g->xf().values().data(), sizeof(double) * d->nx);
    return 0;
}

int corbo_ref_adapt_once(const b200sqp_ocp* d, int n_min, int n_max, double dt_hyst_ratio, const double* x_in, const double* u_in, const double* dt_in,
                         double* x_out, double* u_out, double* dt_out, int32_t* n_out)
{
    return adaptOnce(d, 0, n_min, n_max, dt_hyst_ratio, 0, x_in, u_in, dt_in, x_out, u_out, dt_out, n_out);
}

// the same for adaptGridRedundantControls (non_uniform_finite_differences_variable_grid.cpp:259-352); output capacity n_max + 1 rows
int corbo_ref_adapt_once_redundant(const b200sqp_ocp* d, int n_min, int n_max, int num_backup_nodes, double epsilon, const double* x_in,
                                   const double* u_in, const double* dt_in, double* x_out, double* u_out, double* dt_out, int32_t* n_out)
{
    return adaptOnce(d, 1, n_min, n_max, epsilon, num_backup_nodes, x_in, u_in, dt_in, x_out, u_out, dt_out, n_out);
}

// Time-optimal MPC with grid adaptation, the reference's own classes: a NonUniformFiniteDifferencesVariableGrid with
// setGridAdaptTimeBasedSingleStep(n_max, dt_hyst_ratio) and setNmin(n_min) (non_uniform_finite_differences_variable_grid.cpp:45-50,206-257)
// under the OCP loop of PredictiveController::step (predictive_controller.cpp:66): num_ocp_iterations computes per controller step, the
// first with new_run = true (no adaptation, start state overwritten), the others with new_run = false (adaptGrid before the solve).
// warm_start = 0: the grid re-initialises the trajectories at the adapted size before every solve (non_uniform_full_discretization_grid_base.cpp:90-101).
// x0_seq [steps][nx]: the measured state of every controller step; xref [nx] static reference (goal).
// n_trace [steps*num_ocp_iterations]: grid size N each solve ran on; u0_out [steps][nu]: first control after each step;
// x_last [n_max][nx], u_last [n_max][nu], dt_last [n_max]: the final trajectories (N = last n_trace entry; x has N rows, u and dt N-1).
static int adaptiveSteps(const b200sqp_ocp* d, const b200sqp_lm_options* o, int strategy, int n_min, int n_max, double p0, int p1, int warm_start,
                         int num_ocp_iterations, int steps, const double* x0_seq, const double* xref, int32_t* n_trace, double* u0_out,
                         double* x_last, double* u_last, double* dt_last)
{
    if (d->grid != B200SQP_GRID_FD_NONUNIFORM_VARDT) return -9;
    RefOcp r;
    if (!buildOcp(*d, *o, r)) return -1;
    auto* g = dynamic_cast<NuGridProbe*>(r.grid.get());
    if (!g) return -9;
    selectGridAdaptation(g, strategy, n_min, n_max, p0, p1);
    g->setWarmStart(warm_start != 0);
    StaticReference xr(Eigen::VectorXd(Eigen::Map<const Eigen::VectorXd>(xref, d->nx)));
    ZeroReference uref(d->nu);
    for (int s = 0; s < steps; ++s)
    {
        Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x0_seq + (size_t)s * d->nx, d->nx);
        for (int i = 0; i < num_ocp_iterations; ++i)
        {
            r.ocp->compute(x, xr, uref, nullptr, Time(s * d->dt_ref), i == 0);  // the solver status is not part of this check
            n_trace[(size_t)s * num_ocp_iterations + i] = g->getN();
        }
        Eigen::VectorXd u(d->nu);
        if (!r.ocp->getFirstControlInput(u)) return -3;
        std::memcpy(u0_out + (size_t)s * d->nu, u.data(), sizeof(double) * d->nu);
    }
    const int n = g->getN();
    if (n > n_max + 1) return -10;
    for (int k = 0; k < n - 1; ++k)
    {
        std::memcpy(x_last + (size_t)k * d->nx, g->xs()[k].values().data(), sizeof(double) * d->nx);
        std::memcpy(u_last + (size_t)k * d->nu, g->us()[k].values().data(), sizeof(double) * d->nu);
        dt_last[k] = g->dts()[k].value();
    }
    std::memcpy(x_last + (size_t)(n - 1) * d->nx, g->xf().values().data(), sizeof(double) * d->nx);
    return 0;
}

int corbo_ref_adaptive_steps(const b200sqp_ocp* d, const b200sqp_lm_options* o, int n_min, int n_max, double dt_hyst_ratio, int warm_start,
                             int num_ocp_iterations, int steps, const double* x0_seq, const double* xref, int32_t* n_trace, double* u0_out,
                             double* x_last, double* u_last, double* dt_last)
{
    return adaptiveSteps(d, o, 0, n_min, n_max, dt_hyst_ratio, 0, warm_start, num_ocp_iterations, steps, x0_seq, xref, n_trace, u0_out, x_last, u_last,
                         dt_last);
}

// the same loop with setGridAdaptRedundantControls(n_max, num_backup_nodes, epsilon)
int corbo_ref_adaptive_steps_redundant(const b200sqp_ocp* d, const b200sqp_lm_options* o, int n_min, int n_max, int num_backup_nodes, double epsilon,
                                       int warm_start, int num_ocp_iterations, int steps, const double* x0_seq, const double* xref, int32_t* n_trace,
                                       double* u0_out, double* x_last, double* u_last, double* dt_last)
{
    return adaptiveSteps(d, o, 1, n_min, n_max, epsilon, num_backup_nodes, warm_start, num_ocp_iterations, steps, x0_seq, xref, n_trace, u0_out, x_last,
                         u_last, dt_last);
}

// The reference's known-answer solver tests (optimization/test/test_levenberg_marquardt_sparse.cpp:72-296, excluded from its build)
// run against the compiled reference: SimpleOptimizationProblemWithCallbacks + LevenbergMarquardtSparse, 100 iterations.
// the reference's own getLinearA / getLinearB with its ForwardDifferences (method 0, the default) or CentralDifferences (method 1)
int corbo_ref_linearize(const b200sqp_ocp* d, int method, const double* x0, const double* u0, double* A, double* B)
{
    SystemDynamicsInterface::Ptr dyn = makeDynamics(*d);
    if (!dyn) return -1;
    if (method == 0)
        dyn->setLinearizationMethod(std::make_shared<ForwardDifferences>());
    else
        dyn->setLinearizationMethod(std::make_shared<CentralDifferences>());
    const int nx = d->nx, nu = d->nu;
    Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x0, nx), u = Eigen::Map<const Eigen::VectorXd>(u0, nu);
    if (A)
    {
        Eigen::MatrixXd Am(nx, nx);
        dyn->getLinearA(x, u, Am);
        Eigen::Map<Eigen::MatrixXd>(A, nx, nx) = Am;  // column-major
    }
    if (B)
    {
        Eigen::MatrixXd Bm(nx, nu);
        dyn->getLinearB(x, u, Bm);
        Eigen::Map<Eigen::MatrixXd>(B, nx, nu) = Bm;
    }
    return 0;
}

// the reference's own ForwardDifferences::hessian (method 0) / CentralDifferences::hessian (method 1) templates applied to its
// SystemDynamicsInterface::dynamics as a function of z = [x; u]; H [(nx+nu)^2] column-major
int corbo_ref_dynamics_hessian(const b200sqp_ocp* d, int method, const double* x0, const double* u0, const double* multipliers, double* H)
{
    SystemDynamicsInterface::Ptr dyn = makeDynamics(*d);
    if (!dyn) return -1;
    const int nx = d->nx, nu = d->nu, nz = nx + nu;
    Eigen::VectorXd z(nz);
    z.head(nx) = Eigen::Map<const Eigen::VectorXd>(x0, nx);
    z.tail(nu) = Eigen::Map<const Eigen::VectorXd>(u0, nu);
    auto inc  = [&z](int idx, double inc) { z[idx] += inc; };
    auto eval = [&](Eigen::VectorXd& values) {
        Eigen::VectorXd x = z.head(nx), u = z.tail(nu);
        dyn->dynamics(x, u, values);
    };
    Eigen::MatrixXd Hm(nz, nz);
    if (method == 0)
        ForwardDifferences::hessian(inc, eval, nx, Hm, multipliers);
    else
        CentralDifferences::hessian(inc, eval, nx, Hm, multipliers);
    Eigen::Map<Eigen::MatrixXd>(H, nz, nz) = Hm;
    return 0;
}

int corbo_ref_known_answer(int case_id, int stage, double* x_out, double* expected, double* tol, int32_t* n_out)
{
    SimpleOptimizationProblemWithCallbacks optim;
    LevenbergMarquardtSparse solver;
    solver.setIterations(100);
    auto shifted = [](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) { values[0] = x[0] - 2; };
    int n        = 1;
    double exp3[3] = {0, 0, 0};
    switch (case_id)
    {
        case 0:
            optim.resizeParameterVector(1);
            optim.setX(Eigen::VectorXd::Ones(1));
            optim.setObjectiveFunction(shifted, 1, true);
            exp3[0] = 2, *tol = 1e-6;
            break;
        case 1:
            n = 3;
            optim.resizeParameterVector(3);
            optim.setX(Eigen::VectorXd::Ones(3));
            optim.setObjectiveFunction(
                [](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) {
                    values[0] = x[0] - 5;
                    values[1] = x[1] + 3;
                    values[2] = x[2];
                },
                3, true);
            exp3[0] = 5, exp3[1] = -3, exp3[2] = 0, *tol = 1e-6;
            break;
        case 2:
            n = 2;
            optim.resizeParameterVector(2);
            optim.setX(Eigen::VectorXd::Ones(2));
            optim.setObjectiveFunction(
                [](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) {
                    values[0] = std::sqrt(100) * (x[1] - x[0] * x[0]);
                    values[1] = 1 - x[0];
                },
                2, true);
            exp3[0] = 1, exp3[1] = 1, *tol = 1e-3;
            break;
        case 3:
            optim.resizeParameterVector(1);
            optim.setX(Eigen::VectorXd::Ones(1));
            optim.setObjectiveFunction(shifted, 1, true);
            optim.setEqualityConstraint([](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) { values[0] = x[0] - 3; }, 1);
            solver.setPenaltyWeights(100, 100, 100);
            exp3[0] = 3, *tol = 1e-4;
            break;
        case 4:
            optim.resizeParameterVector(1);
            optim.setX(Eigen::VectorXd::Ones(1));
            optim.setObjectiveFunction(shifted, 1, true);
            optim.setInequalityConstraint([](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) { values[0] = -x[0] + 3; }, 1);
            solver.setPenaltyWeights(100, 100, 100);
            exp3[0] = 3, *tol = 1e-4;
            break;
        case 5:
        case 6:
        {
            optim.resizeParameterVector(1);
            optim.setX(Eigen::VectorXd::Ones(1));
            Eigen::VectorXd b(1);
            b[0] = case_id == 5 ? 5 : -1;
            if (case_id == 5)
                optim.setLowerBounds(b);
            else
                optim.setUpperBounds(b);
            optim.setObjectiveFunction(shifted, 1, true);
            solver.setPenaltyWeights(100, 100, 100);
            exp3[0] = b[0], *tol = 1e-3;
            break;
        }
        case 7:
            n = 2;
            optim.resizeParameterVector(2);
            optim.setLowerBound(0, 2);
            optim.setUpperBound(0, 50);
            optim.setLowerBound(1, -50);
            optim.setUpperBound(1, 50);
            optim.setObjectiveFunction(
                [](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) {
                    values[0] = std::sqrt(0.01) * x[0];
                    values[1] = x[1];
                },
                2, true);
            optim.setInequalityConstraint(
                [](const Eigen::VectorXd& x, Eigen::Ref<Eigen::VectorXd> values) { values[0] = x[1] - 10.0 * x[0] + 10.0; }, 1);
            optim.setParameterValue(0, stage == 0 ? -5 : -1);
            optim.setParameterValue(1, 0);
            if (stage == 1)
            {
                solver.setPenaltyWeights(1, 10, 10);
                solver.setIterations(5000);
            }
            exp3[0] = 2, exp3[1] = 0, *tol = 1e-2;
            break;
        default:
            return -1;
    }
    if (!solver.initialize(&optim)) return -2;
    solver.solve(optim, true, true, nullptr);
    for (int i = 0; i < n; ++i)
    {
        x_out[i]    = optim.getX()[i];
        expected[i] = exp3[i];
    }
    *n_out = n;
    return 0;
}

int corbo_ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
