// Drop-in test of corbo::SolverB200Lm under the UNMODIFIED reference stack:
//   PredictiveController::step -> StructuredOptimalControlProblem::compute -> NlpSolverInterface::solve
// Two identical closed loops (BASELINE.json configs[0]: Van der Pol, FiniteDifferencesGrid N=20, dt=0.1, CN collocation,
// QuadraticFormCost lsq Q=I R=0.1 + QuadraticFinalStateCost, |u|<=1, x0=(1,0.5), 10 LM iterations), one with the reference's
// LevenbergMarquardtSparse and one with SolverB200Lm created through the reference's own factory, must produce the same control
// sequence.  Also exercises the batch front-end.  Built where /root/reference exists (tests/adapter/Makefile), run on the GPU box.
#include <corbo-controllers/predictive_controller.h>
#include <corbo-core/reference_trajectory.h>
#include <corbo-numerics/explicit_integrators.h>
#include <corbo-optimal-control/functions/final_state_constraints.h>
#include <corbo-optimal-control/functions/final_state_cost.h>
#include <corbo-optimal-control/functions/quadratic_cost.h>
#include <corbo-optimal-control/structured_ocp/discretization_grids/finite_differences_grid.h>
#include <corbo-optimal-control/structured_ocp/structured_optimal_control_problem.h>
#include <corbo-optimization/hyper_graph/hyper_graph_optimization_problem_edge_based.h>
#include <corbo-optimization/solver/levenberg_marquardt_sparse.h>
#include <corbo-systems/benchmark/linear_benchmark_systems.h>
#include <corbo-systems/benchmark/nonlinear_benchmark_systems.h>

#include <cmath>
#include <cstdio>
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

int main()
{
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
    for (int variant = 0; variant < 2; ++variant)
    {
        FinalStageConstraint::Ptr fc;
        if (variant == 0)
        {
            Eigen::MatrixXd S = Eigen::MatrixXd::Zero(2, 2);
            S(0, 0) = 2.0;
            S(1, 1) = 0.5;
            fc = std::make_shared<TerminalBall>(S, 0.01);
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
                    variant == 0 ? "TerminalBall" : "TerminalEqualityConstraint", lb.problem->getInequalityDimension(),
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
    std::printf(failures ? "DROP-IN TEST FAILED\n" : "DROP-IN TEST PASSED\n");
    return failures ? 1 : 0;
}
