// Drop-in test of corbo::SolverB200Lm under the UNMODIFIED reference stack:
//   PredictiveController::step -> StructuredOptimalControlProblem::compute -> NlpSolverInterface::solve
// Two identical closed loops (BASELINE.json configs[0]: Van der Pol, FiniteDifferencesGrid N=20, dt=0.1, CN collocation,
// QuadraticFormCost lsq Q=I R=0.1 + QuadraticFinalStateCost, |u|<=1, x0=(1,0.5), 10 LM iterations), one with the reference's
// LevenbergMarquardtSparse and one with SolverB200Lm created through the reference's own factory, must produce the same control
// sequence.  Also exercises the batch front-end.  Built where /root/reference exists (tests/adapter/Makefile), run on the GPU box.
#include <corbo-controllers/predictive_controller.h>
#include <corbo-core/reference_trajectory.h>
#include <corbo-core/time_series.h>
#include <corbo-numerics/explicit_integrators.h>
#include <corbo-optimal-control/functions/final_state_constraints.h>
#include <corbo-optimal-control/functions/final_state_cost.h>
#include <corbo-optimal-control/functions/quadratic_cost.h>
#include <corbo-optimal-control/functions/minimum_time.h>
#include <corbo-optimal-control/structured_ocp/discretization_grids/finite_differences_grid.h>
#include <corbo-optimal-control/structured_ocp/discretization_grids/non_uniform_finite_differences_variable_grid.h>
#include <corbo-optimal-control/structured_ocp/structured_optimal_control_problem.h>
#include <corbo-optimization/hyper_graph/hyper_graph_optimization_problem_edge_based.h>
#include <corbo-optimization/solver/levenberg_marquardt_sparse.h>
#include <corbo-systems/benchmark/linear_benchmark_systems.h>
#include <corbo-systems/benchmark/nonlinear_benchmark_systems.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../control_box_rst_b200/adapter/solver_b200_lm.h"

using namespace corbo;

struct Loop
{
    std::shared_ptr<StructuredOptimalControlProblem> ocp;
    std::shared_ptr<PredictiveController> controller;
    std::shared_ptr<HyperGraphOptimizationProblemEdgeBased> problem;
    SystemDynamicsInterface::Ptr dynamics;
    std::shared_ptr<FiniteDifferencesGrid> grid;
};

static Loop makeLoop(NlpSolverInterface::Ptr solver, int n, FinalStageConstraint::Ptr final_constraint = {},
                     std::shared_ptr<HyperGraphOptimizationProblemEdgeBased> problem = {}, std::shared_ptr<SolverB200Lm> evaluator = {})
{
    Loop l;
    l.dynamics = std::make_shared<VanDerPolOscillator>();
    l.grid     = std::make_shared<FiniteDifferencesGrid>();
    l.grid->setNRef(n);
    l.grid->setDtRef(0.1);
    l.grid->setCostIntegrationRule(FullDiscretizationGridBase::CostIntegrationRule::LeftSum);
    l.problem = problem ? problem : std::make_shared<HyperGraphOptimizationProblemEdgeBased>();
    l.ocp     = std::make_shared<StructuredOptimalControlProblem>(l.grid, l.dynamics, l.problem, solver);
    Eigen::MatrixXd Q = Eigen::MatrixXd::Identity(2, 2), R = Eigen::MatrixXd::Constant(1, 1, 0.1);
    auto stage_cost = std::make_shared<QuadraticFormCost>(Q, R, false, true);
    auto final_cost = std::make_shared<QuadraticFinalStateCost>(Q, true);
    l.ocp->setStageCost(stage_cost);
    l.ocp->setFinalStageCost(final_cost);
    if (final_constraint) l.ocp->setFinalStageConstraint(final_constraint);
    Eigen::VectorXd xlb = Eigen::VectorXd::Constant(2, -CORBO_INF_DBL), xub = Eigen::VectorXd::Constant(2, CORBO_INF_DBL);
    Eigen::VectorXd ulb = Eigen::VectorXd::Constant(1, -1.0), uub = Eigen::VectorXd::Constant(1, 1.0);
    l.ocp->setBounds(xlb, xub, ulb, uub);
    auto b200 = std::dynamic_pointer_cast<SolverB200Lm>(solver);
    if (!b200) b200 = evaluator;  // second surface: the evaluator behind HyperGraphOptimizationProblemB200 needs the same objects
    if (b200)
    {
        // the same objects the OCP got (the functors behind the edges are private in the reference)
        b200->setSystemDynamics(l.dynamics);
        b200->setStageCost(stage_cost);
        b200->setFinalStageCost(final_cost);
        if (final_constraint) b200->setFinalStageConstraint(final_constraint);
    }
    l.controller = std::make_shared<PredictiveController>();
    l.controller->setOptimalControlProblem(l.ocp);
    l.controller->setNumOcpIterations(1);
    return l;
}

// one cold-started OCP solve of a 2-state / 1-input benchmark system under StructuredOptimalControlProblem with `solver`; returns the
// optimised parameter vector (empty on failure)
static Eigen::VectorXd solveBenchmarkSystem(SystemDynamicsInterface::Ptr dynamics, NlpSolverInterface::Ptr solver, const std::vector<double>* parameters,
                                            bool* ok)
{
    auto grid = std::make_shared<FiniteDifferencesGrid>();
    grid->setNRef(20);
    grid->setDtRef(0.1);
    grid->setCostIntegrationRule(FullDiscretizationGridBase::CostIntegrationRule::LeftSum);
    auto problem = std::make_shared<HyperGraphOptimizationProblemEdgeBased>();
    auto ocp     = std::make_shared<StructuredOptimalControlProblem>(grid, dynamics, problem, solver);
    Eigen::MatrixXd Q = Eigen::MatrixXd::Identity(2, 2), R = Eigen::MatrixXd::Constant(1, 1, 0.1);
    auto stage_cost = std::make_shared<QuadraticFormCost>(Q, R, false, true);
    auto final_cost = std::make_shared<QuadraticFinalStateCost>(Q, true);
    ocp->setStageCost(stage_cost);
    ocp->setFinalStageCost(final_cost);
    Eigen::VectorXd xlb = Eigen::VectorXd::Constant(2, -CORBO_INF_DBL), xub = Eigen::VectorXd::Constant(2, CORBO_INF_DBL);
    Eigen::VectorXd ulb = Eigen::VectorXd::Constant(1, -1.5), uub = Eigen::VectorXd::Constant(1, 1.5);
    ocp->setBounds(xlb, xub, ulb, uub);
    if (auto b200 = std::dynamic_pointer_cast<SolverB200Lm>(solver))
    {
        b200->setSystemDynamics(dynamics);
        b200->setStageCost(stage_cost);
        b200->setFinalStageCost(final_cost);
        if (parameters) b200->setSystemDynamicsParameters(*parameters);
    }
    ocp->initialize();
    ZeroReference xref(2), uref(1);
    Eigen::VectorXd x0(2);
    x0 << 1.2, -0.4;
    *ok = ocp->compute(x0, xref, uref, nullptr, Time(0), true);
    Eigen::VectorXd p(problem->getParameterDimension());
    problem->getParameterVector(p);
    return p;
}

// ---- BASELINE.json configs[2]: time-optimal unicycle on the NonUniformFiniteDifferencesVariableGrid (TEB-style: one free dt per interval),
//      MinimumTime in lsq form, goal pose fixed, |v|, |omega| <= 1, dt in [0, 1]; the model class is the product's corbo::Unicycle
struct TebLoop
{
    std::shared_ptr<StructuredOptimalControlProblem> ocp;
    std::shared_ptr<HyperGraphOptimizationProblemEdgeBased> problem;
    SystemDynamicsInterface::Ptr dynamics;
};

static TebLoop makeTebUnicycle(NlpSolverInterface::Ptr solver, int n, bool dt_eq_constraint = false)
{
    TebLoop l;
    l.dynamics = std::make_shared<Unicycle>();
    auto grid  = std::make_shared<NonUniformFiniteDifferencesVariableGrid>();
    grid->setNRef(n);
    grid->setDtRef(0.1);
    grid->setDtBounds(0.0, 1.0);
    grid->disableGridAdaptation();
    grid->setDtEqConstraint(dt_eq_constraint);
    grid->setCostIntegrationRule(NonUniformFullDiscretizationGridBase::CostIntegrationRule::LeftSum);
    Eigen::Matrix<bool, -1, 1> xf_fixed = Eigen::Matrix<bool, -1, 1>::Constant(3, true);
    grid->setXfFixed(xf_fixed);
    l.problem       = std::make_shared<HyperGraphOptimizationProblemEdgeBased>();
    l.ocp           = std::make_shared<StructuredOptimalControlProblem>(grid, l.dynamics, l.problem, solver);
    auto stage_cost = std::make_shared<MinimumTime>(true);
    l.ocp->setStageCost(stage_cost);
    Eigen::VectorXd xlb = Eigen::VectorXd::Constant(3, -CORBO_INF_DBL), xub = Eigen::VectorXd::Constant(3, CORBO_INF_DBL);
    Eigen::VectorXd ulb = Eigen::VectorXd::Constant(2, -1.0), uub = Eigen::VectorXd::Constant(2, 1.0);
    l.ocp->setBounds(xlb, xub, ulb, uub);
    if (auto b200 = std::dynamic_pointer_cast<SolverB200Lm>(solver))
    {
        b200->setSystemDynamics(l.dynamics);
        b200->setStageCost(stage_cost);
    }
    return l;
}

static double relDiff(const Eigen::VectorXd& a, const Eigen::VectorXd& b)
{
    return a.size() == b.size() ? (a - b).cwiseAbs().maxCoeff() / std::max(1.0, a.cwiseAbs().maxCoeff()) : 1e30;
}

static Eigen::VectorXd paramsOf(HyperGraphOptimizationProblemEdgeBased& p)
{
    Eigen::VectorXd v(p.getParameterDimension());
    p.getParameterVector(v);
    return v;
}

// `dropin_test --bench B`: the reference's own API route at batch size B -- B StructuredOptimalControlProblem objects (Van der Pol,
// FiniteDifferencesGrid N=50: BASELINE configs[1]) prepared by the reference's grid update, solved by ONE SolverB200Lm::solveBatch call.
// Timed: the whole solveBatch call on the host clock (hypergraph walks, parameter gather over the vertex objects, H2D, the device
// solve, D2H, parameter scatter back into the vertices).  Prints one JSON object on the last line.
static int benchPlugin(int B, int repetitions)
{
    ZeroReference xref(2), uref(1);
    std::mt19937_64 rng(1235);
    std::uniform_real_distribution<double> dist(-2.0, 2.0);
    std::vector<Loop> loops;
    std::vector<OptimizationProblemInterface*> problems;
    std::vector<Eigen::VectorXd> initial;
    const auto t_build0 = std::chrono::steady_clock::now();
    for (int i = 0; i < B; ++i)
    {
        Eigen::VectorXd x0(2);
        x0 << dist(rng), dist(rng);
        auto dummy = std::make_shared<LevenbergMarquardtSparse>();
        dummy->setIterations(0);  // compute() with zero iterations: grid update + index precomputation only, the initial guess stays
        loops.push_back(makeLoop(dummy, 50));
        loops.back().ocp->initialize();
        loops.back().ocp->compute(x0, xref, uref, nullptr, Time(0), true);
        problems.push_back(loops.back().problem.get());
        initial.push_back(paramsOf(*loops.back().problem));
    }
    const double build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build0).count();
    auto solver = std::make_shared<SolverB200Lm>();
    solver->setIterations(10);
    solver->setSystemDynamics(loops[0].dynamics);
    solver->setStageCost(std::make_shared<QuadraticFormCost>(Eigen::MatrixXd::Identity(2, 2), Eigen::MatrixXd::Constant(1, 1, 0.1), false, true));
    solver->setFinalStageCost(std::make_shared<QuadraticFinalStateCost>(Eigen::MatrixXd::Identity(2, 2), true));
    double best = 1e30, total = 0;
    for (int r = 0; r < repetitions + 1; ++r)
    {
        for (int i = 0; i < B; ++i) problems[i]->setParameterVector(initial[i]);  // cold start again (not timed)
        const auto t0 = std::chrono::steady_clock::now();
        if (!solver->solveBatch(problems, true, nullptr, nullptr))
        {
            std::printf("{\"error\": \"solveBatch: %s\"}\n", solver->lastError().c_str());
            return 1;
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r > 0)  // the first call creates the device handle and runs the structure self-check
        {
            best = std::min(best, s);
            total += s;
        }
    }
    const double mean = total / repetitions;
    std::printf("{\"objects\": %d, \"n_grid\": 50, \"iterations\": 10, \"repetitions\": %d, \"solve_batch_ms_mean\": %.4f, \"solve_batch_ms_best\": %.4f, "
                "\"value\": %.6g, \"unit\": \"iters/s\", \"kernel_ms\": %.4f, \"build_objects_s\": %.3f}\n",
                B, repetitions, 1e3 * mean, 1e3 * best, B * 10.0 / mean, solver->lastSolveMilliseconds(), build_s);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 3 && std::strcmp(argv[1], "--bench") == 0) return benchPlugin(std::atoi(argv[2]), argc >= 4 ? std::atoi(argv[3]) : 5);
    int failures = 0;
    // ---- 1. factory registration: the plugin is found by name like every reference solver -------------------------------------
    NlpSolverInterface::Ptr from_factory = NlpSolverFactory::instance().create("SolverB200Lm");
    if (!from_factory || !std::dynamic_pointer_cast<SolverB200Lm>(from_factory))
    {
        std::printf("FAIL: SolverB200Lm not registered in Factory<NlpSolverInterface>\n");
        return 1;
    }
    std::printf("ok: Factory<NlpSolverInterface> creates SolverB200Lm\n");

    // ---- 2. closed loop under PredictiveController --------------------------------------------------------------------------------
    auto lm_ref = std::make_shared<LevenbergMarquardtSparse>();
    lm_ref->setIterations(10);
    auto lm_b200 = std::dynamic_pointer_cast<SolverB200Lm>(from_factory);
    lm_b200->setIterations(10);
    Loop a = makeLoop(lm_ref, 20), b = makeLoop(lm_b200, 20);
    ZeroReference xref(2), uref(1);
    IntegratorExplicitRungeKutta4 rk4;
    Eigen::VectorXd xa(2), xb(2);
    xa << 1.0, 0.5;
    xb = xa;
    double worst_u = 0, worst_x = 0;
    for (int s = 0; s < 15; ++s)
    {
        TimeSeries::Ptr ua = std::make_shared<TimeSeries>(), ub = std::make_shared<TimeSeries>();
        TimeSeries::Ptr sa = std::make_shared<TimeSeries>(), sb = std::make_shared<TimeSeries>();
        bool oka = a.controller->step(xa, xref, uref, Duration(0.1), Time(0.1 * s), ua, sa);
        bool okb = b.controller->step(xb, xref, uref, Duration(0.1), Time(0.1 * s), ub, sb);
        if (!oka || !okb)
        {
            std::printf("FAIL: controller step %d returned %d / %d (%s)\n", s, (int)oka, (int)okb, lm_b200->lastError().c_str());
            return 1;
        }
        Eigen::VectorXd u_a(1), u_b(1);
        a.ocp->getFirstControlInput(u_a);
        b.ocp->getFirstControlInput(u_b);
        worst_u = std::max(worst_u, std::abs(u_a[0] - u_b[0]));
        Eigen::VectorXd na(2), nb(2);
        rk4.solveIVP(xa, u_a, 0.1, *a.dynamics, na);
        rk4.solveIVP(xb, u_b, 0.1, *b.dynamics, nb);
        xa = na;
        xb = nb;
        worst_x = std::max(worst_x, (xa - xb).cwiseAbs().maxCoeff());
        if (s < 3 || s == 14) std::printf("  step %2d  u_ref=% .9f  u_b200=% .9f   obj_ref=%.9g obj_b200=%.9g\n", s, u_a[0], u_b[0], a.ocp->getCurrentObjectiveValue(), b.ocp->getCurrentObjectiveValue());
    }
    std::printf("closed loop (15 MPC steps): max |u_ref - u_b200| = %.3e, max |x_ref - x_b200| = %.3e\n", worst_u, worst_x);
    if (!(worst_u <= 2e-6 && worst_x <= 2e-6))
    {
        std::printf("FAIL: closed loops diverge\n");
        ++failures;
    }

    // ---- 3. batch front-end: 64 OCP objects of one structure, one device call, against per-object reference solves ----------------
    const int B = 64;
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> dist(-2.0, 2.0);
    std::vector<Loop> loops_ref, loops_b200;
    std::vector<OptimizationProblemInterface*> problems;
    auto batch_solver = std::make_shared<SolverB200Lm>();
    batch_solver->setIterations(10);
    double worst = 0;
    std::vector<Eigen::VectorXd> x0s;
    for (int i = 0; i < B; ++i)
    {
        Eigen::VectorXd x0(2);
        x0 << dist(rng), dist(rng);
        x0s.push_back(x0);
        auto solver_i = std::make_shared<LevenbergMarquardtSparse>();
        solver_i->setIterations(10);
        loops_ref.push_back(makeLoop(solver_i, 50));
        loops_ref.back().ocp->initialize();
        loops_ref.back().ocp->compute(x0, xref, uref, nullptr, Time(0), true);
        // the batched side: let the reference build grid + hypergraph (grid update, index precomputation), then solve all at once
        auto dummy = std::make_shared<LevenbergMarquardtSparse>();
        dummy->setIterations(0);  // compute() with zero iterations only prepares the structure and leaves the initial guess
        loops_b200.push_back(makeLoop(dummy, 50));
        loops_b200.back().ocp->initialize();
        loops_b200.back().ocp->compute(x0, xref, uref, nullptr, Time(0), true);
        problems.push_back(loops_b200.back().problem.get());
    }
    batch_solver->setSystemDynamics(loops_b200[0].dynamics);
    batch_solver->setStageCost(std::make_shared<QuadraticFormCost>(Eigen::MatrixXd::Identity(2, 2), Eigen::MatrixXd::Constant(1, 1, 0.1), false, true));
    batch_solver->setFinalStageCost(std::make_shared<QuadraticFinalStateCost>(Eigen::MatrixXd::Identity(2, 2), true));
    std::vector<SolverStatus> statuses;
    std::vector<double> objs;
    if (!batch_solver->solveBatch(problems, true, &statuses, &objs))
    {
        std::printf("FAIL: solveBatch: %s\n", batch_solver->lastError().c_str());
        return 1;
    }
    for (int i = 0; i < B; ++i)
    {
        Eigen::VectorXd pr(loops_ref[i].problem->getParameterDimension()), pb(pr.size());
        loops_ref[i].problem->getParameterVector(pr);
        loops_b200[i].problem->getParameterVector(pb);
        worst = std::max(worst, (pr - pb).cwiseAbs().maxCoeff() / std::max(1.0, pr.cwiseAbs().maxCoeff()));
    }
    std::printf("batch front-end (%d OCP objects, N=50): max relative trajectory difference vs reference = %.3e, kernel %.3f ms\n", B, worst,
                batch_solver->lastSolveMilliseconds());
    if (!(worst <= 1e-5))
    {
        std::printf("FAIL: batched trajectories differ\n");
        ++failures;
    }

    // ---- 3b. final-stage constraints: TerminalBall (the inequality edge with its active-set rows) and TerminalEqualityConstraint ----
    for (int variant = 0; variant < 3; ++variant)
    {
        FinalStageConstraint::Ptr fc;
        if (variant == 0)
        {
            Eigen::MatrixXd S = Eigen::MatrixXd::Zero(2, 2);
            S(0, 0) = 2.0;
            S(1, 1) = 0.5;
            fc = std::make_shared<TerminalBall>(S, 0.01);
        }
        else if (variant == 2)
        {
            // TerminalBallInheritFromCost: S is taken from the quadratic final cost (here Qf = I), gamma is its own member
            auto inherit    = std::make_shared<TerminalBallInheritFromCost>();
            inherit->_gamma = 0.02;
            fc              = inherit;
        }
        else
        {
            Eigen::VectorXd tx(2);
            tx << 0.1, -0.05;
            fc = std::make_shared<TerminalEqualityConstraint>(tx);
        }
        Loop lr = makeLoop(std::make_shared<LevenbergMarquardtSparse>(), 30, fc);
        Loop lb = makeLoop(std::make_shared<SolverB200Lm>(), 30, fc);
        lr.ocp->initialize();
        lb.ocp->initialize();
        Eigen::VectorXd x0(2);
        x0 << 1.5, -0.7;
        bool ok_r = lr.ocp->compute(x0, xref, uref, nullptr, Time(0), true);
        bool ok_b = lb.ocp->compute(x0, xref, uref, nullptr, Time(0), true);
        Eigen::VectorXd pr(lr.problem->getParameterDimension()), pb(lb.problem->getParameterDimension());
        lr.problem->getParameterVector(pr);
        lb.problem->getParameterVector(pb);
        double diff = pr.size() == pb.size() ? (pr - pb).cwiseAbs().maxCoeff() / std::max(1.0, pr.cwiseAbs().maxCoeff()) : 1e30;
        std::printf("final-stage constraint %s: ineq dim %d, eq dim %d, max relative trajectory difference vs reference = %.3e\n",
                    variant == 0 ? "TerminalBall" : (variant == 2 ? "TerminalBallInheritFromCost" : "TerminalEqualityConstraint"), lb.problem->getInequalityDimension(),
                    lb.problem->getEqualityDimension(), diff);
        if (!ok_r || !ok_b || !(diff <= 1e-5))
        {
            std::printf("FAIL: final-stage constraint variant %d (ok_ref=%d ok_b200=%d)\n", variant, (int)ok_r, (int)ok_b);
            ++failures;
        }
    }

    // ---- 4. error behaviour: structures outside the registry -> SolverStatus::Error, never a CPU fallback ---------------------------
    {
        auto s = std::make_shared<SolverB200Lm>();  // no setSystemDynamics
        Loop l = makeLoop(std::make_shared<LevenbergMarquardtSparse>(), 10);
        l.ocp->initialize();
        Eigen::VectorXd x0(2);
        x0 << 1.0, 0.5;
        l.ocp->compute(x0, xref, uref, nullptr, Time(0), true);
        double obj = 0;
        SolverStatus st = s->solve(*l.problem, true, true, &obj);
        if (st != SolverStatus::Error || obj != -1)
        {
            std::printf("FAIL: missing functors must yield SolverStatus::Error and obj_value -1\n");
            ++failures;
        }
        else
            std::printf("ok: unsupported structure -> SolverStatus::Error (%s)\n", s->lastError().c_str());
    }
    // ---- 5. second surface: the reference's OWN LevenbergMarquardtSparse on a HyperGraphOptimizationProblemB200 (created by name through
    //         the reference's factory), i.e. its Jacobians come from the device; against the same solver on the stock problem class.
    //         Van der Pol is polynomial: Jacobian values, pattern and parameter drift are bit-identical, so the iterates must be too.
    {
        auto from_problem_factory = HyperGraphOptimizationProblemFactory::instance().create("HyperGraphOptimizationProblemB200");
        auto device_problem       = std::dynamic_pointer_cast<HyperGraphOptimizationProblemB200>(from_problem_factory);
        if (!device_problem)
        {
            std::printf("FAIL: HyperGraphOptimizationProblemB200 not registered in Factory<BaseHyperGraphOptimizationProblem>\n");
            return 1;
        }
        auto evaluator = std::make_shared<SolverB200Lm>();
        device_problem->setDeviceEvaluator(evaluator);
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>(), s_dev = std::make_shared<LevenbergMarquardtSparse>();
        s_ref->setIterations(10);
        s_dev->setIterations(10);
        Loop lr = makeLoop(s_ref, 30);
        Loop ld = makeLoop(s_dev, 30, {}, device_problem, evaluator);
        lr.ocp->initialize();
        ld.ocp->initialize();
        Eigen::VectorXd x0(2);
        x0 << -1.2, 0.8;
        double worst_p = 0;
        bool ok5 = true;
        for (int s = 0; s < 3; ++s)  // cold start, then two warm-started solves with a moved start state
        {
            ok5 = lr.ocp->compute(x0, xref, uref, nullptr, Time(0.1 * s), true) && ok5;
            ok5 = ld.ocp->compute(x0, xref, uref, nullptr, Time(0.1 * s), true) && ok5;
            Eigen::VectorXd pr(lr.problem->getParameterDimension()), pd(ld.problem->getParameterDimension());
            lr.problem->getParameterVector(pr);
            ld.problem->getParameterVector(pd);
            worst_p = std::max(worst_p, pr.size() == pd.size() ? (pr - pd).cwiseAbs().maxCoeff() : 1e30);
            Eigen::VectorXd u(1);
            lr.ocp->getFirstControlInput(u);
            Eigen::VectorXd xn(2);
            rk4.solveIVP(x0, u, 0.1, *lr.dynamics, xn);
            x0 = xn;
        }
        std::printf("second surface: reference LevenbergMarquardtSparse on device Jacobians (%d device evaluations): max |p_ref - p_dev| = %.3e, "
                    "objective %.12g vs %.12g\n",
                    device_problem->deviceJacobianEvaluations(), worst_p, lr.ocp->getCurrentObjectiveValue(), ld.ocp->getCurrentObjectiveValue());
        if (!ok5 || device_problem->failed() || device_problem->deviceJacobianEvaluations() < 3 || worst_p != 0.0)
        {
            std::printf("FAIL: second surface (ok=%d failed=%d, %s)\n", (int)ok5, (int)device_problem->failed(), evaluator->lastError().c_str());
            ++failures;
        }
    }
    // ---- 6. the other benchmark systems behind the plugin: a parameter-free class is recognised by its type, a class with setters but no
    //         getters (nonlinear_benchmark_systems.h) needs setSystemDynamicsParameters(); a wrong or missing value is an Error, never a
    //         silently different model
    {
        auto newRef = [] { auto s = std::make_shared<LevenbergMarquardtSparse>(); s->setIterations(10); return s; };
        auto newDev = [] { auto s = std::make_shared<SolverB200Lm>(); s->setIterations(10); return s; };
        auto makeDuffing = [] { auto d = std::make_shared<DuffingOscillator>(); d->setParameters(1.0, -1.0, 1.0); return d; };
        const std::vector<double> duffing_parameters{1.0, -1.0, 1.0}, wrong_parameters{1.0, -1.0, 2.0};
        bool ok_r = false, ok_d = false;
        Eigen::VectorXd p_r = solveBenchmarkSystem(makeDuffing(), newRef(), nullptr, &ok_r);
        Eigen::VectorXd p_d = solveBenchmarkSystem(makeDuffing(), newDev(), &duffing_parameters, &ok_d);
        double err = (ok_r && ok_d && p_r.size() == p_d.size()) ? (p_r - p_d).cwiseAbs().maxCoeff() : 1e30;
        std::printf("Duffing oscillator behind the plugin: max |p_ref - p_b200| = %.3e\n", err);
        if (!(err <= 2e-5)) { std::printf("FAIL: Duffing solve differs\n"); ++failures; }
        ok_r = ok_d = false;
        p_r = solveBenchmarkSystem(std::make_shared<ArtsteinsCircle>(), newRef(), nullptr, &ok_r);
        p_d = solveBenchmarkSystem(std::make_shared<ArtsteinsCircle>(), newDev(), nullptr, &ok_d);
        err = (ok_r && ok_d && p_r.size() == p_d.size()) ? (p_r - p_d).cwiseAbs().maxCoeff() : 1e30;
        std::printf("Artstein's circle behind the plugin (recognised by type): max |p_ref - p_b200| = %.3e\n", err);
        if (!(err <= 2e-5)) { std::printf("FAIL: Artstein's circle solve differs\n"); ++failures; }
        auto makeLinear = [] {
            auto d = std::make_shared<LinearStateSpaceModel>();
            Eigen::MatrixXd A(2, 2), Bm(2, 1);
            A << -0.3, 1.1, -2.0, -0.5;
            Bm << 0.2, 1.5;
            d->setParameters(A, Bm);
            return d;
        };
        const std::vector<double> linear_parameters{-0.3, -2.0, 1.1, -0.5, 0.2, 1.5};  // A column-major, then B
        ok_r = ok_d = false;
        p_r = solveBenchmarkSystem(makeLinear(), newRef(), nullptr, &ok_r);
        p_d = solveBenchmarkSystem(makeLinear(), newDev(), &linear_parameters, &ok_d);
        err = (ok_r && ok_d && p_r.size() == p_d.size()) ? (p_r - p_d).cwiseAbs().maxCoeff() : 1e30;
        std::printf("LinearStateSpaceModel (2x2 A, 2x1 B) behind the plugin: max |p_ref - p_b200| = %.3e\n", err);
        if (!(err <= 2e-5)) { std::printf("FAIL: linear state-space solve differs\n"); ++failures; }
        auto dev_wrong = newDev(), dev_missing = newDev();
        bool ok_w = true, ok_m = true;
        solveBenchmarkSystem(makeDuffing(), dev_wrong, &wrong_parameters, &ok_w);
        solveBenchmarkSystem(makeDuffing(), dev_missing, nullptr, &ok_m);
        if (ok_w || ok_m)
        {
            std::printf("FAIL: wrong (%d) / missing (%d) dynamics parameters must fail the solve\n", (int)ok_w, (int)ok_m);
            ++failures;
        }
        else
            std::printf("ok: wrong parameters -> Error (%s); missing parameters -> Error (%s)\n", dev_wrong->lastError().c_str(),
                        dev_missing->lastError().c_str());
    }
    // ---- 7. BASELINE configs[2] through the plugin: time-optimal unicycle (corbo::Unicycle, b200_systems.h) on the non-uniform grid.
    //         Two consecutive solves per object (the live dt vertices change after every solve and must not count as a new structure:
    //         the device handle survives, and with new_run = false the penalty weights are ADAPTED like LevenbergMarquardtSparse's), then
    //         a batch of 8 objects solved twice with one device call each (every instance then has its own dt values).
    {
        ZeroReference xref3(3), uref2(2);
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>();
        auto s_dev = std::make_shared<SolverB200Lm>();
        s_ref->setIterations(5);
        s_dev->setIterations(5);
        s_ref->setWeightAdapation(2, 2, 2, 50, 50, 50);
        s_dev->setWeightAdapation(2, 2, 2, 50, 50, 50);
        TebLoop lr = makeTebUnicycle(s_ref, 20), lb = makeTebUnicycle(s_dev, 20);
        lr.ocp->initialize();
        lb.ocp->initialize();
        Eigen::VectorXd x0(3);
        x0 << -1.5, 0.8, 0.6;
        double worst7 = 0;
        bool ok7      = true;
        for (int s = 0; s < 3; ++s)
        {
            const bool new_run = s != 1;  // the second solve continues the run: weights adapted, not reset
            ok7 = lr.ocp->compute(x0, xref3, uref2, nullptr, Time(0.1 * s), new_run) && ok7;
            ok7 = lb.ocp->compute(x0, xref3, uref2, nullptr, Time(0.1 * s), new_run) && ok7;
            worst7 = std::max(worst7, relDiff(paramsOf(*lr.problem), paramsOf(*lb.problem)));
            x0[0] += 0.05;
            x0[2] -= 0.02;
        }
        std::printf("time-optimal unicycle (non-uniform grid, N=20), 3 consecutive solves incl. one with adapted weights: max relative "
                    "trajectory difference vs reference = %.3e (objective %.9g vs %.9g)\n",
                    worst7, lr.ocp->getCurrentObjectiveValue(), lb.ocp->getCurrentObjectiveValue());
        if (!ok7 || !(worst7 <= 1e-4))
        {
            std::printf("FAIL: time-optimal unicycle through the plugin (ok=%d, %s)\n", (int)ok7, s_dev->lastError().c_str());
            ++failures;
        }
        const int Bt = 8;
        std::vector<TebLoop> refs, devs;
        std::vector<OptimizationProblemInterface*> probs;
        std::mt19937_64 rng7(77);
        std::uniform_real_distribution<double> pos(-2.0, 2.0), ang(-3.0, 3.0);
        std::vector<Eigen::VectorXd> starts;
        for (int i = 0; i < Bt; ++i)
        {
            Eigen::VectorXd xs(3);
            xs << pos(rng7), pos(rng7), ang(rng7);
            starts.push_back(xs);
            auto si = std::make_shared<LevenbergMarquardtSparse>();
            si->setIterations(5);
            refs.push_back(makeTebUnicycle(si, 20));
            refs.back().ocp->initialize();
            auto dummy = std::make_shared<LevenbergMarquardtSparse>();
            dummy->setIterations(0);
            devs.push_back(makeTebUnicycle(dummy, 20));
            devs.back().ocp->initialize();
            probs.push_back(devs.back().problem.get());
        }
        auto batch7 = std::make_shared<SolverB200Lm>();
        batch7->setIterations(5);
        batch7->setSystemDynamics(devs[0].dynamics);
        batch7->setStageCost(std::make_shared<MinimumTime>(true));
        double worst7b = 0;
        bool ok7b      = true;
        for (int round = 0; round < 2; ++round)
        {
            for (int i = 0; i < Bt; ++i)
            {
                ok7b = refs[i].ocp->compute(starts[i], xref3, uref2, nullptr, Time(0.1 * round), true) && ok7b;
                ok7b = devs[i].ocp->compute(starts[i], xref3, uref2, nullptr, Time(0.1 * round), true) && ok7b;  // grid update only (0 iterations)
            }
            if (!batch7->solveBatch(probs, true, nullptr, nullptr))
            {
                std::printf("FAIL: solveBatch on the non-uniform grid, round %d: %s\n", round, batch7->lastError().c_str());
                ok7b = false;
                break;
            }
            for (int i = 0; i < Bt; ++i)
            {
                worst7b = std::max(worst7b, relDiff(paramsOf(*refs[i].problem), paramsOf(*devs[i].problem)));
                starts[i][0] += 0.03;
            }
        }
        std::printf("time-optimal unicycle, %d objects x 2 rounds through solveBatch: max relative trajectory difference vs reference = %.3e\n", Bt, worst7b);
        if (!ok7b || !(worst7b <= 1e-4))
        {
            std::printf("FAIL: batched time-optimal unicycle\n");
            ++failures;
        }
    }
    // ---- 8. BASELINE configs[4] through the plugin: 12-state quadrotor (corbo::Quadrotor, parameters through its getters), FD grid N=12
    {
        auto build = [](NlpSolverInterface::Ptr solver, std::shared_ptr<HyperGraphOptimizationProblemEdgeBased>& problem) {
            auto dynamics = std::make_shared<Quadrotor>(1.2, 9.81, 0.012, 0.011, 0.021);
            auto grid     = std::make_shared<FiniteDifferencesGrid>();
            grid->setNRef(12);
            grid->setDtRef(0.05);
            grid->setCostIntegrationRule(FullDiscretizationGridBase::CostIntegrationRule::LeftSum);
            problem  = std::make_shared<HyperGraphOptimizationProblemEdgeBased>();
            auto ocp = std::make_shared<StructuredOptimalControlProblem>(grid, dynamics, problem, solver);
            Eigen::MatrixXd Q = Eigen::MatrixXd::Identity(12, 12), R = 0.1 * Eigen::MatrixXd::Identity(4, 4);
            auto stage_cost = std::make_shared<QuadraticFormCost>(Q, R, false, true);
            auto final_cost = std::make_shared<QuadraticFinalStateCost>(Q, true);
            ocp->setStageCost(stage_cost);
            ocp->setFinalStageCost(final_cost);
            Eigen::VectorXd xlb = Eigen::VectorXd::Constant(12, -CORBO_INF_DBL), xub = Eigen::VectorXd::Constant(12, CORBO_INF_DBL);
            Eigen::VectorXd ulb(4), uub(4);
            ulb << 0.0, -1.0, -1.0, -1.0;
            uub << 2.0 * 1.2 * 9.81, 1.0, 1.0, 1.0;
            ocp->setBounds(xlb, xub, ulb, uub);
            if (auto b200 = std::dynamic_pointer_cast<SolverB200Lm>(solver))
            {
                b200->setSystemDynamics(dynamics);
                b200->setStageCost(stage_cost);
                b200->setFinalStageCost(final_cost);
            }
            ocp->initialize();
            return ocp;
        };
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>();
        auto s_dev = std::make_shared<SolverB200Lm>();
        s_ref->setIterations(6);
        s_dev->setIterations(6);
        std::shared_ptr<HyperGraphOptimizationProblemEdgeBased> pr, pd;
        auto ocp_r = build(s_ref, pr), ocp_d = build(s_dev, pd);
        ZeroReference xref12(12), uref4(4);
        Eigen::VectorXd x0 = Eigen::VectorXd::Zero(12);
        x0[0] = 0.6;
        x0[1] = -0.4;
        x0[2] = 0.3;
        x0[3] = 0.1;
        x0[4] = -0.15;
        x0[5] = 0.05;
        const bool ok_r = ocp_r->compute(x0, xref12, uref4, nullptr, Time(0), true);
        const bool ok_d = ocp_d->compute(x0, xref12, uref4, nullptr, Time(0), true);
        const double diff = relDiff(paramsOf(*pr), paramsOf(*pd));
        std::printf("quadrotor (12 states, FD grid N=12) through the plugin: max relative trajectory difference vs reference = %.3e "
                    "(objective %.9g vs %.9g)\n", diff, ocp_r->getCurrentObjectiveValue(), ocp_d->getCurrentObjectiveValue());
        if (!ok_r || !ok_d || !(diff <= 1e-4))
        {
            std::printf("FAIL: quadrotor through the plugin (ok_ref=%d ok_b200=%d, %s)\n", (int)ok_r, (int)ok_d, s_dev->lastError().c_str());
            ++failures;
        }
        if (!SystemDynamicsFactory::instance().create("Unicycle") || !SystemDynamicsFactory::instance().create("Quadrotor"))
        {
            std::printf("FAIL: corbo::Unicycle / corbo::Quadrotor not registered in Factory<SystemDynamicsInterface>\n");
            ++failures;
        }
    }
    // ---- 9. a time-varying state reference: the reference's own DiscreteTimeReferenceTrajectory (a time series, linearly interpolated)
    //         handed to compute() and -- the same kind of object -- to the plugin through setStateReference(); three MPC steps with
    //         the reference window moving along the series
    {
        auto makeReference = [] {
            auto series = std::make_shared<TimeSeries>();
            for (int k = 0; k < 60; ++k)
            {
                Eigen::VectorXd v(2);
                v << 0.5 * std::sin(0.15 * k), 0.02 * k;
                series->add(0.1 * k, v);
            }
            return std::make_shared<DiscreteTimeReferenceTrajectory>(series, TimeSeries::Interpolation::Linear);
        };
        auto xr_ref = makeReference(), xr_dev = makeReference();
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>();
        auto s_dev = std::make_shared<SolverB200Lm>();
        s_ref->setIterations(8);
        s_dev->setIterations(8);
        Loop lr = makeLoop(s_ref, 25), ld = makeLoop(s_dev, 25);
        s_dev->setStateReference(xr_dev);
        lr.ocp->initialize();
        ld.ocp->initialize();
        ZeroReference uref1(1);
        Eigen::VectorXd x0(2);
        x0 << 0.3, -0.2;
        double worst9 = 0;
        bool ok9      = true;
        for (int s = 0; s < 3; ++s)
        {
            ok9 = lr.ocp->compute(x0, *xr_ref, uref1, nullptr, Time(0.1 * s), true) && ok9;
            ok9 = ld.ocp->compute(x0, *xr_dev, uref1, nullptr, Time(0.1 * s), true) && ok9;
            worst9 = std::max(worst9, relDiff(paramsOf(*lr.problem), paramsOf(*ld.problem)));
            Eigen::VectorXd u(1), xn(2);
            lr.ocp->getFirstControlInput(u);
            IntegratorExplicitRungeKutta4 rk;
            rk.solveIVP(x0, u, 0.1, *lr.dynamics, xn);
            x0 = xn;
        }
        std::printf("time-varying reference (DiscreteTimeReferenceTrajectory, 3 MPC steps): max relative trajectory difference vs reference = %.3e "
                    "(objective %.9g vs %.9g)\n", worst9, lr.ocp->getCurrentObjectiveValue(), ld.ocp->getCurrentObjectiveValue());
        if (!ok9 || !(worst9 <= 1e-5))
        {
            std::printf("FAIL: time-varying reference through the plugin (ok=%d, %s)\n", (int)ok9, s_dev->lastError().c_str());
            ++failures;
        }
    }
    // ---- 10. full (non-diagonal) weight matrices: QuadraticFormCost / QuadraticFinalStateCost with the upper Cholesky square root and a
    //          non-zero static reference (the reference's zero-reference branch for a non-diagonal Q is broken, quadratic_cost.cpp:112)
    {
        auto build = [](NlpSolverInterface::Ptr solver, std::shared_ptr<HyperGraphOptimizationProblemEdgeBased>& problem) {
            auto dynamics = std::make_shared<VanDerPolOscillator>();
            auto grid     = std::make_shared<FiniteDifferencesGrid>();
            grid->setNRef(20);
            grid->setDtRef(0.1);
            grid->setCostIntegrationRule(FullDiscretizationGridBase::CostIntegrationRule::LeftSum);
            problem  = std::make_shared<HyperGraphOptimizationProblemEdgeBased>();
            auto ocp = std::make_shared<StructuredOptimalControlProblem>(grid, dynamics, problem, solver);
            Eigen::MatrixXd Q(2, 2), Qf(2, 2), R = Eigen::MatrixXd::Constant(1, 1, 0.1);
            Q << 2.0, 0.3, 0.3, 1.0;
            Qf << 3.0, -0.4, -0.4, 2.0;
            auto stage_cost = std::make_shared<QuadraticFormCost>(Q, R, false, true);
            auto final_cost = std::make_shared<QuadraticFinalStateCost>(Qf, true);
            ocp->setStageCost(stage_cost);
            ocp->setFinalStageCost(final_cost);
            Eigen::VectorXd xlb = Eigen::VectorXd::Constant(2, -CORBO_INF_DBL), xub = Eigen::VectorXd::Constant(2, CORBO_INF_DBL);
            Eigen::VectorXd ulb = Eigen::VectorXd::Constant(1, -1.0), uub = Eigen::VectorXd::Constant(1, 1.0);
            ocp->setBounds(xlb, xub, ulb, uub);
            if (auto b200 = std::dynamic_pointer_cast<SolverB200Lm>(solver))
            {
                b200->setSystemDynamics(dynamics);
                b200->setStageCost(stage_cost);
                b200->setFinalStageCost(final_cost);
            }
            ocp->initialize();
            return ocp;
        };
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>();
        auto s_dev = std::make_shared<SolverB200Lm>();
        s_ref->setIterations(8);
        s_dev->setIterations(8);
        Eigen::VectorXd goal(2);
        goal << 0.2, -0.1;
        auto xr_ref = std::make_shared<StaticReference>(goal), xr_dev = std::make_shared<StaticReference>(goal);
        s_dev->setStateReference(xr_dev);
        std::shared_ptr<HyperGraphOptimizationProblemEdgeBased> pr, pd;
        auto ocp_r = build(s_ref, pr), ocp_d = build(s_dev, pd);
        ZeroReference uref1(1);
        Eigen::VectorXd x0(2);
        x0 << 1.1, -0.6;
        const bool ok_r = ocp_r->compute(x0, *xr_ref, uref1, nullptr, Time(0), true);
        const bool ok_d = ocp_d->compute(x0, *xr_dev, uref1, nullptr, Time(0), true);
        const double diff = relDiff(paramsOf(*pr), paramsOf(*pd));
        std::printf("full weight matrices Q, Qf (upper Cholesky square roots): max relative trajectory difference vs reference = %.3e "
                    "(objective %.9g vs %.9g)\n", diff, ocp_r->getCurrentObjectiveValue(), ocp_d->getCurrentObjectiveValue());
        if (!ok_r || !ok_d || !(diff <= 1e-5))
        {
            std::printf("FAIL: full weight matrices through the plugin (ok_ref=%d ok_b200=%d, %s)\n", (int)ok_r, (int)ok_d, s_dev->lastError().c_str());
            ++failures;
        }
    }
    // ---- 11. the time-optimal unicycle with setDtEqConstraint(true): TwoScalarEqualEdges between consecutive dt vertices
    {
        ZeroReference xref3(3), uref2(2);
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>();
        auto s_dev = std::make_shared<SolverB200Lm>();
        s_ref->setIterations(6);
        s_dev->setIterations(6);
        TebLoop lr = makeTebUnicycle(s_ref, 16, true), lb = makeTebUnicycle(s_dev, 16, true);
        lr.ocp->initialize();
        lb.ocp->initialize();
        Eigen::VectorXd x0(3);
        x0 << 1.2, -0.9, 0.4;
        double worst11 = 0;
        bool ok11      = true;
        for (int s = 0; s < 2; ++s)
        {
            ok11 = lr.ocp->compute(x0, xref3, uref2, nullptr, Time(0.1 * s), true) && ok11;
            ok11 = lb.ocp->compute(x0, xref3, uref2, nullptr, Time(0.1 * s), true) && ok11;
            worst11 = std::max(worst11, relDiff(paramsOf(*lr.problem), paramsOf(*lb.problem)));
            x0[1] += 0.04;
        }
        std::printf("time-optimal unicycle with dt equality edges (eq dim %d): max relative trajectory difference vs reference = %.3e "
                    "(objective %.9g vs %.9g)\n", lb.problem->getEqualityDimension(), worst11, lr.ocp->getCurrentObjectiveValue(),
                    lb.ocp->getCurrentObjectiveValue());
        if (!ok11 || !(worst11 <= 1e-4))
        {
            std::printf("FAIL: dt equality edges through the plugin (ok=%d, %s)\n", (int)ok11, s_dev->lastError().c_str());
            ++failures;
        }
    }
    // ---- 12. grid adaptation: the reference's own NonUniformFiniteDifferencesVariableGrid adapts the grid
    //      (setGridAdaptTimeBasedSingleStep) between the OCP iterations of a controller step; the plugin follows the changing structure
    //      with the penalty weights carried over, and solveBatch buckets objects of different grid sizes
    {
        ZeroReference xref3(3), uref2(2);
        auto adaptive = [](NlpSolverInterface::Ptr solver) {
            TebLoop l = makeTebUnicycle(solver, 14);
            auto* grid = dynamic_cast<NonUniformFiniteDifferencesVariableGrid*>(l.ocp->getDiscretizationGrid().get());
            grid->setGridAdaptTimeBasedSingleStep(30, 0.1);
            grid->setNmin(3);
            grid->setWarmStart(true);
            l.ocp->initialize();
            return l;
        };
        // (a) one OCP with the plugin as its solver, 3 controller steps x 3 OCP iterations
        auto s_ref = std::make_shared<LevenbergMarquardtSparse>();
        auto s_dev = std::make_shared<SolverB200Lm>();
        s_ref->setIterations(6);
        s_dev->setIterations(6);
        TebLoop lr = adaptive(s_ref), lb = adaptive(s_dev);
        Eigen::VectorXd x0(3);
        x0 << 4.6, -1.2, 0.5;
        double worst12 = 0;
        bool ok12      = true;
        int n_first = 1000, n_last = 0;  // smallest / largest grid seen
        for (int s = 0; s < 3; ++s)
        {
            for (int it = 0; it < 3; ++it)
            {
                ok12 = lr.ocp->compute(x0, xref3, uref2, nullptr, Time(0.1 * s), it == 0) && ok12;
                ok12 = lb.ocp->compute(x0, xref3, uref2, nullptr, Time(0.1 * s), it == 0) && ok12;
                const int nr = lr.ocp->getDiscretizationGrid()->getN(), nb = lb.ocp->getDiscretizationGrid()->getN();
                n_first = std::min(n_first, nr), n_last = std::max(n_last, nr);
                if (nr != nb)
                {
                    std::printf("FAIL: grid sizes diverge (step %d, OCP iteration %d): %d vs %d\n", s, it, nr, nb);
                    ok12 = false;
                    break;
                }
                worst12 = std::max(worst12, relDiff(paramsOf(*lr.problem), paramsOf(*lb.problem)));
            }
            x0[0] -= 0.05;
        }
        std::printf("adaptive time-optimal unicycle through the plugin: grids of %d..%d points over 9 solves, max relative trajectory difference "
                    "vs reference = %.3e\n", n_first, n_last, worst12);
        // chained solves of a trigonometric model: the bar of tests/test_gpu_grid_adaptation.py
        if (!ok12 || n_first == n_last || !(worst12 <= 5e-4))
        {
            std::printf("FAIL: grid adaptation through the plugin (ok=%d, %s)\n", (int)ok12, s_dev->lastError().c_str());
            ++failures;
        }
        // (b) six objects whose grids drift apart, one solveBatch per OCP iteration
        const int Ba = 6;
        const double start_x[Ba] = {0.4, 0.9, 1.5, 2.2, 3.0, 3.8};
        std::vector<TebLoop> refs, devs;
        std::vector<OptimizationProblemInterface*> probs;
        std::vector<Eigen::VectorXd> starts;
        for (int i = 0; i < Ba; ++i)
        {
            Eigen::VectorXd xs(3);
            xs << start_x[i], 0.3 * (i % 3) - 0.3, 0.2 * i - 0.4;
            starts.push_back(xs);
            auto si = std::make_shared<LevenbergMarquardtSparse>();
            si->setIterations(6);
            refs.push_back(adaptive(si));
            auto dummy = std::make_shared<LevenbergMarquardtSparse>();
            dummy->setIterations(0);
            devs.push_back(adaptive(dummy));
            probs.push_back(devs.back().problem.get());
        }
        auto batch12 = std::make_shared<SolverB200Lm>();
        batch12->setIterations(6);
        batch12->setSystemDynamics(devs[0].dynamics);
        batch12->setStageCost(std::make_shared<MinimumTime>(true));
        double worst12b = 0;
        bool ok12b      = true;
        int n_min_seen = 1000, n_max_seen = 0;
        for (int s = 0; s < 2 && ok12b; ++s)
            for (int it = 0; it < 3 && ok12b; ++it)
            {
                for (int i = 0; i < Ba; ++i)
                {
                    ok12b = refs[i].ocp->compute(starts[i], xref3, uref2, nullptr, Time(0.1 * s), it == 0) && ok12b;
                    ok12b = devs[i].ocp->compute(starts[i], xref3, uref2, nullptr, Time(0.1 * s), it == 0) && ok12b;  // grid update + adaptation only
                }
                if (!batch12->solveBatch(probs, it == 0, nullptr, nullptr))
                {
                    std::printf("FAIL: solveBatch over mixed grid sizes (step %d, OCP iteration %d): %s\n", s, it, batch12->lastError().c_str());
                    ok12b = false;
                    break;
                }
                for (int i = 0; i < Ba; ++i)
                {
                    const int nr = refs[i].ocp->getDiscretizationGrid()->getN(), nb = devs[i].ocp->getDiscretizationGrid()->getN();
                    n_min_seen = std::min(n_min_seen, nb), n_max_seen = std::max(n_max_seen, nb);
                    if (nr != nb)
                    {
                        std::printf("FAIL: object %d: grid sizes diverge (%d vs %d)\n", i, nr, nb);
                        ok12b = false;
                    }
                    else
                        worst12b = std::max(worst12b, relDiff(paramsOf(*refs[i].problem), paramsOf(*devs[i].problem)));
                }
            }
        std::printf("adaptive time-optimal unicycle, %d objects x 6 solves through solveBatch, grid sizes %d..%d in one batch: max relative "
                    "trajectory difference vs reference = %.3e\n", Ba, n_min_seen, n_max_seen, worst12b);
        if (!ok12b || n_min_seen == n_max_seen || !(worst12b <= 5e-4))
        {
            std::printf("FAIL: batched grid adaptation through the plugin\n");
            ++failures;
        }
    }
    std::printf(failures ? "DROP-IN TEST FAILED\n" : "DROP-IN TEST PASSED\n");
    return failures ? 1 : 0;
}
