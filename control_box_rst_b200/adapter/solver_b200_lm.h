// corbo::SolverB200Lm -- the reference-side plugin of this project: a corbo::NlpSolverInterface
// (src/optimization/include/corbo-optimization/solver/nlp_solver_interface.h:67-118) that keeps the interface and parameters of
// corbo::LevenbergMarquardtSparse (solver/levenberg_marquardt_sparse.h:68-163) and runs the whole LM/SQP inner loop on a B200
// through the C ABI of libb200sqp.so (include/b200sqp.h).  It drops in under
// StructuredOptimalControlProblem::initialize/compute/reset (src/optimal_control/src/structured_ocp/
// structured_optimal_control_problem.cpp:61,134,204) and PredictiveController::step (src/controllers/src/predictive_controller.cpp:66)
// without touching reference sources.
//
// Discovery gap: a solver only receives OptimizationProblemInterface&.  The hypergraph walk yields the grid
// kind and size, dimensions, bounds, fixed masks, x0 and the parameter vector; the functors behind the edges are private in the
// reference, so the objects the user handed to the OCP are handed to this solver as well (setSystemDynamics, setCollocation /
// setIntegrator, setStageCost, setFinalStageCost, setFinalStageConstraint, setStateReference).  Anything outside the closed registry makes solve() return
// SolverStatus::Error -- there is no CPU fallback.  After every structure upload the device residual vector is compared with the
// reference's own computeValues on the host to catch a mis-extraction.
#ifndef CONTROL_BOX_RST_B200_ADAPTER_SOLVER_B200_LM_H_
#define CONTROL_BOX_RST_B200_ADAPTER_SOLVER_B200_LM_H_

#include <corbo-core/reference_trajectory.h>
#include <corbo-numerics/finite_differences_collocation.h>
#include <corbo-numerics/integrator_interface.h>
#include <corbo-optimal-control/functions/stage_functions.h>
#include <corbo-optimization/hyper_graph/hyper_graph_optimization_problem_edge_based.h>
#include <corbo-optimization/solver/nlp_solver_interface.h>
#include <corbo-systems/system_dynamics_interface.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200sqp.h"
#include "b200_systems.h"  // corbo::Unicycle, corbo::Quadrotor: the models of BASELINE configs[2] / configs[4] for the reference's object model

namespace corbo {

class SolverB200Lm : public NlpSolverInterface
{
 public:
    using Ptr = std::shared_ptr<SolverB200Lm>;

    SolverB200Lm();
    ~SolverB200Lm() override;

    // ---- NlpSolverInterface ----------------------------------------------------------------------------------------------
    NlpSolverInterface::Ptr getInstance() const override { return std::make_shared<SolverB200Lm>(); }
    bool isLsqSolver() const override { return true; }
    bool initialize(OptimizationProblemInterface* problem = nullptr) override;
    SolverStatus solve(OptimizationProblemInterface& problem, bool new_structure = true, bool new_run = true, double* obj_value = nullptr) override;
    void clear() override;

    // ---- LevenbergMarquardtSparse parameters (levenberg_marquardt_sparse.h:85-90) -------------------------------------------------
    void setIterations(int iterations) { _opts.iterations = iterations; }
    void setPenaltyWeights(double weight_eq, double weight_ineq, double weight_bounds);
    void setWeightAdapation(double factor_eq, double factor_ineq, double factor_bounds, double max_eq, double max_ineq, double max_bounds);

    // ---- the objects behind the edges (same shared_ptrs the user gave to StructuredOptimalControlProblem / the grid) -----------
    void setSystemDynamics(SystemDynamicsInterface::Ptr dynamics) { _dynamics = dynamics; }
    // parameters of a dynamics class that has setters but no getters in the reference (DuffingOscillator: damping, spring_alpha,
    // spring_beta; SimplePendulum: m, l, g, rho; MasslessPendulum: omega0; ToyExample: mu -- nonlinear_benchmark_systems.h; a LinearStateSpaceModel
    // of 2x1, 3x1, 4x1 or 4x2 states x inputs: A column-major then B column-major -- linear_benchmark_systems.h:216): the values
    // the user passed to setParameters().  A mismatch is caught by the residual self-check after every structure upload.
    void setSystemDynamicsParameters(const std::vector<double>& parameters) { _dynamics_parameters = parameters; }
    void setCollocation(FiniteDifferencesCollocationInterface::Ptr collocation) { _collocation = collocation; }
    void setIntegrator(NumericalIntegratorExplicitInterface::Ptr integrator) { _integrator = integrator; }
    void setStageCost(StageCost::Ptr stage_cost) { _stage_cost = stage_cost; }
    void setFinalStageCost(FinalStageCost::Ptr final_cost) { _final_cost = final_cost; }
    void setFinalStageConstraint(FinalStageConstraint::Ptr final_constraint) { _final_constraint = final_constraint; }
    void setStateReference(ReferenceTrajectoryInterface::Ptr xref) { _xref = xref; }
    void setDevice(int device) { _device = device; }

    // ---- batch front-end: B OCP objects, one device call per structure (SURVEY.md section 7 "hard parts") ---------------------------
    // problems[i] hold the initial parameters before and the optimised ones after the call; statuses/obj_values may be null.
    // The objects are OCPs of one kind; their grids may differ in size (time-optimal grids after the reference's own adaptGrid,
    // SURVEY.md section 8f row 2): the batch is then bucketed by structure and every bucket keeps its own device handle.
    bool solveBatch(const std::vector<OptimizationProblemInterface*>& problems, bool new_run, std::vector<SolverStatus>* statuses,
                    std::vector<double>* obj_values);

    // ---- second plugin surface: the residual vector and the combined sparse Jacobian of `problem` at its current
    // parameters, evaluated on the device (b200sqp_evaluate) -- what LevenbergMarquardtSparse::computeValues (:222-246) and
    // ...EdgeBased::computeCombinedSparseJacobian (:1480-1753) deliver, for solvers other than this one.  Like the reference's Jacobian
    // evaluation it leaves the round-trip drift of the in-place differences in the problem's parameters.  values / jacobian may be null.
    bool evaluateOnDevice(OptimizationProblemInterface& problem, double weight_eq, double weight_ineq, double weight_bounds, Eigen::VectorXd* values,
                          Eigen::SparseMatrix<double>* jacobian);

    const std::string& lastError() const { return _error; }
    double lastSolveMilliseconds() const;

 private:
    bool solveUniform(const std::vector<OptimizationProblemInterface*>& problems, const b200sqp_lm_options& opts, std::vector<SolverStatus>* statuses,
                      std::vector<double>* obj_values);
    bool describe(OptimizationProblemInterface& problem, b200sqp_ocp& ocp, std::vector<double>& x0, std::vector<double>& xref);
    bool upload(OptimizationProblemInterface& problem, int batch, std::vector<double>* x0_out = nullptr, std::vector<double>* xref_out = nullptr);
    bool instanceData(OptimizationProblemInterface& problem, double* x0, double* xref, std::string& error) const;
    bool uploadReferenceTrajectory(int batch);
    bool selfCheck(OptimizationProblemInterface& problem);
    SolverStatus fail(const std::string& msg);

    b200sqp_lm_options _opts;
    b200sqp_handle _handle = nullptr;
    b200sqp_ocp _ocp;
    b200sqp_dims _dims;
    int _batch  = 0;
    bool _fresh = false;  // device state was (re)built by the last upload()
    int _device = 0;
    std::string _error;

    SystemDynamicsInterface::Ptr _dynamics;
    std::vector<double> _dynamics_parameters;
    FiniteDifferencesCollocationInterface::Ptr _collocation;
    NumericalIntegratorExplicitInterface::Ptr _integrator;
    StageCost::Ptr _stage_cost;
    FinalStageCost::Ptr _final_cost;
    FinalStageConstraint::Ptr _final_constraint;
    ReferenceTrajectoryInterface::Ptr _xref;
    std::vector<int32_t> _col_ptr, _row_idx;  // CSC pattern of the combined Jacobian of the uploaded structure
    // penalty weights live in the solver object like the reference's (_weight_eq, ...; levenberg_marquardt_sparse.cpp:83-86, 264-287):
    // they survive a change of structure (a grid that was adapted between two solves of one run)
    double _w_eq = 0, _w_ineq = 0, _w_bounds = 0;
    bool _weights_initialised = false;
    std::vector<double> _x0_buf, _xref_buf, _params_buf, _chi2_buf;  // host staging of a batch, reused across calls
    std::vector<int32_t> _status_buf;
    std::map<int, std::shared_ptr<SolverB200Lm>> _by_size;  // buckets of a batch with mixed grid sizes, keyed by parameter dimension
};

FACTORY_REGISTER_NLP_SOLVER(SolverB200Lm)

// The second surface as a reference-side class: a hypergraph optimisation problem whose combined sparse Jacobian comes from the
// device, so that ANY least-squares solver of the reference (e.g. its own LevenbergMarquardtSparse) runs on device Jacobians.
// Registered like the reference's problem classes (hyper_graph_optimization_problem_base.h:266).  Everything else (values,
// increments, backups, indices) is inherited: only the 91 % of an iteration the survey measured in computeCombinedSparseJacobian move.
// The evaluator is a SolverB200Lm that was handed the functors (setSystemDynamics, setStageCost, ...); calls the device cannot serve
// (partial category selections, structures outside the registry) are an error, not a silent host evaluation.
class HyperGraphOptimizationProblemB200 : public HyperGraphOptimizationProblemEdgeBased
{
 public:
    BaseHyperGraphOptimizationProblem::Ptr getInstance() const override { return std::make_shared<HyperGraphOptimizationProblemB200>(); }
    void setDeviceEvaluator(std::shared_ptr<SolverB200Lm> evaluator) { _evaluator = evaluator; }
    int deviceJacobianEvaluations() const { return _device_evaluations; }
    bool failed() const { return _failed; }

    void computeCombinedSparseJacobian(Eigen::SparseMatrix<double>& jacobian, bool objective_lsq, bool equality, bool inequality,
                                       bool finite_combined_bounds, bool active_ineq = false, double weight_eq = 1.0, double weight_ineq = 1.0,
                                       double weight_bounds = 1.0, const Eigen::VectorXd* values = nullptr,
                                       const Eigen::VectorXi* col_nnz = nullptr) override;

 private:
    std::shared_ptr<SolverB200Lm> _evaluator;
    int _device_evaluations = 0;
    bool _failed            = false;
};

FACTORY_REGISTER_HYPER_GRAPH_OPTIMIZATION_PROBLEM(HyperGraphOptimizationProblemB200)

}  // namespace corbo

#endif  // CONTROL_BOX_RST_B200_ADAPTER_SOLVER_B200_LM_H_
