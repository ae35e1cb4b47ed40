/*
 * b200sqp.h -- C ABI of libb200sqp.so: the batched, device-resident Levenberg-Marquardt / SQP inner loop for
 * control_box_rst's hypergraph-structured direct-transcription OCPs on NVIDIA B200 (sm_100a).
 *
 * The reference (rst-tu-dortmund/control_box_rst) has no C ABI: its plugin boundary is the C++ class
 * corbo::NlpSolverInterface (src/optimization/include/corbo-optimization/solver/nlp_solver_interface.h:67-118).
 * The entry points below are what a corbo::NlpSolverInterface subclass binds to (see
 * control_box_rst_b200/adapter/solver_b200_lm.{h,cpp} and INTEGRATION.md); each one names the reference interface it
 * replaces.  Plain pointers and sizes only; no C++/torch types.  All functions return 0 on success or a negative
 * b200sqp_error; they never throw.  Host buffers are owned by the caller, device buffers by the library.  A handle owns
 * one CUDA stream and is not thread-safe.
 */
#ifndef B200SQP_H_
#define B200SQP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SQP_MAX_NX 16
#define B200SQP_MAX_NU 8
#define B200SQP_MAX_DYN_PARAMS 32

typedef enum {
    B200SQP_OK                = 0,
    B200SQP_ERR_INVALID       = -1, /* bad argument / null pointer / inconsistent descriptor */
    B200SQP_ERR_UNSUPPORTED   = -2, /* structure outside the closed functor registry -> corbo::SolverStatus::Error, no CPU fallback */
    B200SQP_ERR_CUDA          = -3, /* CUDA runtime failure (b200sqp_last_error() has the text) */
    B200SQP_ERR_NO_DEVICE     = -4, /* no sm_100 class device visible */
    B200SQP_ERR_NOT_LSQ       = -5  /* problem is not in least-squares form (levenberg_marquardt_sparse.cpp:50-54) */
} b200sqp_error;

/* corbo::SolverStatus (src/optimization/include/corbo-optimization/types.h:30) */
typedef enum { B200SQP_STATUS_CONVERGED = 0, B200SQP_STATUS_EARLY_TERMINATED = 1, B200SQP_STATUS_INFEASIBLE = 2, B200SQP_STATUS_ERROR = 3 } b200sqp_status;

/* System dynamics registry: corbo::SystemDynamicsInterface::dynamics (src/systems/include/corbo-systems/system_dynamics_interface.h:121).
 * 0..4 and 7..10 restate models of src/systems/include/corbo-systems/benchmark/nonlinear_benchmark_systems.h and
 * linear_benchmark_systems.h (every model of the nonlinear header is in the registry);
 * UNICYCLE and QUADROTOR do not exist in the reference and are defined by this project on both sides. */
typedef enum {
    B200SQP_DYN_VAN_DER_POL      = 0, /* nonlinear_benchmark_systems.h:52-60, params[0] = a */
    B200SQP_DYN_DUFFING          = 1, /* nonlinear_benchmark_systems.h DuffingOscillator */
    B200SQP_DYN_SIMPLE_PENDULUM  = 2, /* nonlinear_benchmark_systems.h SimplePendulum, params = m,l,g,rho */
    B200SQP_DYN_CART_POLE        = 3, /* nonlinear_benchmark_systems.h:337-352; params = mc,mp,l,g are the class' private constants (1, 0.3, 0.5, 9.81): pass those or zeros, other values are refused */
    B200SQP_DYN_DOUBLE_INTEGRATOR = 4, /* linear_benchmark_systems.h DoubleIntegratorDiscreteTime's continuous twin: x'' = u */
    B200SQP_DYN_UNICYCLE         = 5, /* new: x' = v cos(th), y' = v sin(th), th' = w */
    B200SQP_DYN_QUADROTOR        = 6, /* new: 12-state rigid-body quadrotor, params = m,g,Ixx,Iyy,Izz */
    B200SQP_DYN_FREE_SPACE_ROCKET = 7, /* nonlinear_benchmark_systems.h:174-183 FreeSpaceRocket (s, v, m), no parameters */
    B200SQP_DYN_MASSLESS_PENDULUM = 8, /* nonlinear_benchmark_systems.h:281-290 MasslessPendulum, params[0] = omega0 */
    B200SQP_DYN_TOY_EXAMPLE      = 9, /* nonlinear_benchmark_systems.h:426-436 ToyExample, params[0] = mu */
    B200SQP_DYN_ARTSTEINS_CIRCLE = 10, /* nonlinear_benchmark_systems.h:483-492 ArtsteinsCircle, no parameters */
    B200SQP_DYN_LINEAR_2X1       = 11, /* linear_benchmark_systems.h:186-214 LinearStateSpaceModel, f = A x + B u, with a 2x2 A and a
                                          2x1 B: params = A column-major (nx*nx values), then B column-major (nx*nu values) */
    B200SQP_DYN_LINEAR_3X1       = 12, /* the same with a 3x3 A and a 3x1 B */
    B200SQP_DYN_LINEAR_4X1       = 13, /* the same with a 4x4 A and a 4x1 B */
    B200SQP_DYN_LINEAR_4X2       = 14, /* the same with a 4x4 A and a 4x2 B */
    B200SQP_DYN_TRIPLE_INTEGRATOR = 15, /* linear_benchmark_systems.h:71-82 SerialIntegratorSystem(3): x''' = u / T, params[0] = T */
    B200SQP_DYN_QUAD_INTEGRATOR  = 16  /* SerialIntegratorSystem(4), params[0] = T */
} b200sqp_dynamics;

/* Discretization grids (vertex sets + edge factories), src/optimal_control/.../discretization_grids/ */
typedef enum {
    B200SQP_GRID_FD_UNIFORM           = 0, /* FiniteDifferencesGrid: fixed x0, single fixed dt (finite_differences_grid.cpp:38-154) */
    B200SQP_GRID_FD_NONUNIFORM_VARDT  = 1, /* NonUniformFiniteDifferencesVariableGrid: one free dt per interval (non_uniform_finite_differences_variable_grid.cpp:60) */
    B200SQP_GRID_MULTIPLE_SHOOTING    = 2  /* MultipleShootingGrid, one control per interval (multiple_shooting_grid.cpp:38-197) */
} b200sqp_grid;

/* FiniteDifferencesCollocationInterface (src/numerics/include/corbo-numerics/finite_differences_collocation.h:119-241) */
typedef enum { B200SQP_COLL_FORWARD = 0, B200SQP_COLL_BACKWARD = 1, B200SQP_COLL_MIDPOINT = 2, B200SQP_COLL_CRANK_NICOLSON = 3 } b200sqp_collocation;

/* NumericalIntegratorExplicitInterface (src/numerics/include/corbo-numerics/explicit_integrators.h) */
typedef enum { B200SQP_INT_EULER = 0, B200SQP_INT_RK4 = 1 } b200sqp_integrator;

/* Stage cost registry (src/optimal_control/src/functions/) */
typedef enum {
    B200SQP_COST_NONE           = 0,
    B200SQP_COST_QUADRATIC_LSQ  = 1, /* QuadraticFormCost(Q,R,integral=false,lsq=true), diagonal Q/R (quadratic_cost.cpp:100-184) */
    B200SQP_COST_MINIMUM_TIME_LSQ = 2 /* MinimumTime(lsq=true) (minimum_time.h:49-78) */
} b200sqp_stage_cost;

/* Final-stage constraints (src/optimal_control/include/corbo-optimal-control/functions/final_state_constraints.h) */
typedef enum {
    B200SQP_FINAL_CONSTRAINT_NONE     = 0,
    B200SQP_FINAL_CONSTRAINT_EQUALITY = 1, /* TerminalEqualityConstraint: nx equality rows */
    B200SQP_FINAL_CONSTRAINT_BALL     = 2  /* TerminalBall, diagonal S: one inequality row (active-set rule of
                                              hyper_graph_optimization_problem_base.cpp:278-289, ..._edge_based.cpp:1565-1616) */
} b200sqp_final_constraint;

/*
 * One OCP structure shared by all instances of a batch.  It carries exactly what the hypergraph walk of a
 * StructuredOptimalControlProblem yields (SURVEY.md section 8b "discovery gap"): grid kind and size, functor ids and their
 * parameters, bounds, fixed masks.  Per-instance data (x0, references, the parameter vector) is set separately.
 */
typedef struct b200sqp_ocp {
    int32_t grid;         /* b200sqp_grid */
    int32_t dynamics;     /* b200sqp_dynamics */
    int32_t collocation;  /* b200sqp_collocation (FD grids) */
    int32_t integrator;   /* b200sqp_integrator (shooting grids) */
    int32_t n_grid;       /* N: number of grid points (N-1 intervals), >= 2 */
    int32_t nx, nu;       /* must match the dynamics id */
    int32_t stage_cost;   /* b200sqp_stage_cost */
    int32_t final_cost;   /* 0 none, 1 QuadraticFinalStateCost(Qf, lsq=true) diagonal (final_state_cost.cpp:73-90) */
    int32_t zero_x_ref;   /* ReferenceTrajectoryInterface::isZero() of xref (quadratic_cost.cpp:107) */
    int32_t zero_u_ref;   /* must be 1 with lsq control cost (the reference's non-zero-uref lsq branch returns a scalar, quadratic_cost.cpp:161) */
    int32_t xf_fixed[B200SQP_MAX_NX]; /* FullDiscretizationGridBase::setXfFixed */
    double dt_ref;        /* grid dt (fixed) or initial dt (variable grids) */
    double dt_lb, dt_ub;  /* NonUniformFiniteDifferencesVariableGrid::setDtBounds; ignored for fixed-dt grids */
    double dyn_params[B200SQP_MAX_DYN_PARAMS];
    double q_diag[B200SQP_MAX_NX];  /* diagonal of Q  (stage state cost)   */
    double r_diag[B200SQP_MAX_NU];  /* diagonal of R  (stage control cost) */
    double qf_diag[B200SQP_MAX_NX]; /* diagonal of Qf (final state cost)   */
    double x_lb[B200SQP_MAX_NX], x_ub[B200SQP_MAX_NX]; /* |bound| >= 2e30 (CORBO_INF_DBL, core/types.h:53) means unbounded */
    double u_lb[B200SQP_MAX_NU], u_ub[B200SQP_MAX_NU];
    /* Final-stage constraint on the (not fully fixed) final state, one edge added after all interval edges
     * (finite_differences_grid.cpp:131-144, nlp_functions.cpp:204-218): */
    int32_t final_constraint;            /* b200sqp_final_constraint */
    double term_xref[B200SQP_MAX_NX];    /* TerminalEqualityConstraint::setXRef (final_state_constraints.h:167-232): rows x_N - term_xref */
    double term_s_diag[B200SQP_MAX_NX];  /* TerminalBall::setWeightS, diagonal S (final_state_constraints.h:38-107) */
    double term_gamma;                   /* TerminalBall::setGamma: one inequality row (x_N - xref)^T S (x_N - xref) - gamma <= 0 */
    /* Full (non-diagonal) weight matrices, row-major, used instead of q_diag / r_diag / qf_diag when the flag is 1.  Like the reference
     * (QuadraticFormCost::setWeightQ / setWeightR, quadratic_cost.cpp:32-76; QuadraticFinalStateCost::setWeightQf, final_state_cost.cpp:38-58)
     * a matrix that is diagonal to 1e-10 falls back to the element-wise square root of its diagonal; otherwise the lsq residual is
     * U (x - xref) with the UPPER Cholesky factor U, Q = U^T U (Eigen::LLT<MatrixXd, Upper>::matrixU()).  A non-diagonal Q needs a non-zero
     * state reference: with a zero reference the reference's lsq branch returns the scalar x^T U x into a vector (quadratic_cost.cpp:112),
     * which is not reproduced (B200SQP_ERR_UNSUPPORTED). */
    int32_t q_dense, r_dense, qf_dense;
    double q_full[B200SQP_MAX_NX * B200SQP_MAX_NX];
    double r_full[B200SQP_MAX_NU * B200SQP_MAX_NU];
    double qf_full[B200SQP_MAX_NX * B200SQP_MAX_NX];
    /* NonUniformFiniteDifferencesVariableGrid::setDtEqConstraint (non_uniform_finite_differences_variable_grid.cpp:150-154): one
     * TwoScalarEqualEdge (edges/misc_edges.h:40-67) per pair of consecutive intervals, value dt_k - dt_{k-1}, created right after the
     * dynamics edge of interval k (k >= 1) in the equality category.  Couples consecutive dt vertices. */
    int32_t dt_eq_constraint;
} b200sqp_ocp;

/* LevenbergMarquardtSparse parameters (levenberg_marquardt_sparse.h:85-90,112-124); defaults 10 / 2,2,2 / 1,1,1 / 500,500,500 */
typedef struct b200sqp_lm_options {
    int32_t iterations;
    double weight_eq, weight_ineq, weight_bounds;                      /* setPenaltyWeights */
    double adapt_factor_eq, adapt_factor_ineq, adapt_factor_bounds;    /* setWeightAdapation */
    double adapt_max_eq, adapt_max_ineq, adapt_max_bounds;
} b200sqp_lm_options;

/* Dimensions the reference reports through OptimizationProblemInterface (levenberg_marquardt_sparse.cpp:56-71) */
typedef struct b200sqp_dims {
    int32_t n_params;    /* getParameterDimension() */
    int32_t m_lsq;       /* getLsqObjectiveDimension() */
    int32_t m_eq;        /* getEqualityDimension() */
    int32_t m_ineq;      /* getInequalityDimension() */
    int32_t m_bounds;    /* finiteCombinedBoundsDimension() */
    int32_t nnz_jacobian;      /* computeSparseJacobian*NNZ() summed (structural, explicit zeros included) */
    int32_t nnz_hessian_upper; /* structural nnz of triu(J^T J) */
    int32_t n_blocks, block_dim; /* block-tridiagonal view used on the device */
    int64_t algorithmic_bytes_per_iteration; /* s*[2*(nnzJ+nnzH+nnzL+2m+2n)+4n], s=8 */
} b200sqp_dims;

typedef struct b200sqp_solver* b200sqp_handle;

/* ---- lifetime ------------------------------------------------------------------------------------------------------ */
/* NlpSolverInterface::initialize + the new_structure branch of LevenbergMarquardtSparse::solve (levenberg_marquardt_sparse.cpp:48-80):
 * derives dimensions and index maps, allocates all device state for `batch` instances on CUDA device `device`. */
int b200sqp_create(const b200sqp_ocp* ocp, int32_t batch, int32_t device, b200sqp_handle* out);
/* NlpSolverInterface::clear + destruction */
int b200sqp_destroy(b200sqp_handle h);
const char* b200sqp_last_error(void);
/* 1 if the library was built with the CUDA kernels and a usable device is present, else 0 (never a CPU fallback) */
int b200sqp_device_available(void);

/* ---- structure / indexing (bit-exact with the reference) --------------------------------------------------------------- */
/* Host-only, needs no GPU: usable on a handle-less descriptor. */
int b200sqp_dims_of(const b200sqp_ocp* ocp, b200sqp_dims* out);
int b200sqp_get_dims(b200sqp_handle h, b200sqp_dims* out);
/* VertexSetInterface::computeVertexIndices (vertex_set.cpp:405-418) over FullDiscretizationGridBase::computeActiveVertices
 * (full_discretization_grid_base.cpp:514-527): for every grid point k, the parameter index of x_k[0] / u_k[0] / dt_k, or -1 if fixed. */
int b200sqp_vertex_indices(const b200sqp_ocp* ocp, int32_t* x_idx /*[N]*/, int32_t* u_idx /*[N-1]*/, int32_t* dt_idx /*[N-1]*/);
/* OptimizationEdgeSet::computeEdgeIndices (edge_set.cpp:31-42,101-166): row offset inside its category of the state-cost,
 * control-cost and dynamics edge of every interval; final-cost edge index in final_cost_idx (or -1). */
int b200sqp_edge_indices(const b200sqp_ocp* ocp, int32_t* state_cost_idx /*[N-1]*/, int32_t* control_cost_idx /*[N-1]*/,
                         int32_t* dt_cost_idx /*[2*(N-1)]*/, int32_t* dynamics_idx /*[N-1]*/, int32_t* final_cost_idx /*[1]*/);
/* row offset inside the equality category of the TwoScalarEqualEdge between dt_{k-1} and dt_k, k = 1..N-2 (entry 0 is -1); all -1
 * without dt_eq_constraint */
int b200sqp_dt_equality_indices(const b200sqp_ocp* ocp, int32_t* dt_eq_idx /*[N-1]*/);
/* row offset of the final-stage constraint edge inside the equality (eq_idx) or inequality (ineq_idx) category, -1 if absent */
int b200sqp_final_constraint_indices(const b200sqp_ocp* ocp, int32_t* eq_idx /*[1]*/, int32_t* ineq_idx /*[1]*/);
/* CSC pattern of computeCombinedSparseJacobian (hyper_graph_optimization_problem_edge_based.cpp:1480-1753), rows lsq->eq->ineq->bounds */
int b200sqp_jacobian_pattern(const b200sqp_ocp* ocp, int32_t* col_ptr /*[n+1]*/, int32_t* row_idx /*[nnzJ]*/);

/* ---- per-instance data ------------------------------------------------------------------------------------------------- */
/* x0: [batch*nx] measured start states (FullDiscretizationGridBase::update :101); xref: [batch*nx] static state reference or NULL
 * (then zero / xf = 0).  Host pointers.  Asynchronous on the handle's stream: pageable buffers are staged before the call returns,
 * PINNED buffers (cudaHostAlloc / cudaHostRegister) are read in place by a kernel and must stay valid and unchanged until
 * b200sqp_synchronize or the next blocking call on the handle (b200sqp_solve / _step / _mpc_step / _get_*). */
int b200sqp_set_problem_data(b200sqp_handle h, const double* x0, const double* xref);
/* Time-varying state reference (a ReferenceTrajectoryInterface with isStatic() == false, src/core/include/corbo-core/reference_trajectory.h:60-95):
 * xref_traj [batch][n_grid][nx], row k = what getReferenceCached(k) returns for that instance.  The cost edge of grid point k then measures
 * x_k against row k (quadratic_cost.cpp:116-121, final_state_cost.cpp); the initial guess of b200sqp_initialize_trajectories becomes the
 * reference trajectory itself (full_discretization_grid_base.cpp:181-228); goal, fixed goal components and the final-stage constraint use
 * the last row.  Call after b200sqp_set_problem_data (whose xref argument, also NULL, re-establishes a static reference); NULL here does the
 * same.  Full-discretisation grids only: B200SQP_ERR_UNSUPPORTED on the shooting grid, whose cold start in the reference ignores the
 * measured state for non-static references (shooting_grid_base.cpp:259,278).  Host pointer, copied before the call returns. */
int b200sqp_set_reference_trajectory(b200sqp_handle h, const double* xref_traj);
/* FullDiscretizationGridBase::initializeSequences (full_discretization_grid_base.cpp:134-179): linear x0 -> xref interpolation,
 * u = 0, dt = dt_ref, on the device. */
int b200sqp_initialize_trajectories(b200sqp_handle h);
/* Parameter vectors in the reference's own order (param index = vertex_idx + free component), [batch*n_params], host pointers. */
int b200sqp_set_params(b200sqp_handle h, const double* params);
int b200sqp_get_params(b200sqp_handle h, double* params);
/* FullDiscretizationGridBase::getFirstControlInput: u_0 of every instance, [batch*nu] */
int b200sqp_get_first_controls(b200sqp_handle h, double* u0);

/* Moving-horizon warm start between two MPC steps, on the device: for every instance
 * FullDiscretizationGridBase::findNearestState + warmStartShifting (full_discretization_grid_base.cpp:230-318) against the new
 * measurement x0_new [batch*nx] (host pointer), then the start state is replaced by the measurement (:101) and fixed goal components
 * by the reference (:102-106).  num_shift [batch] (host, may be NULL) receives the shift each instance applied.
 * FiniteDifferencesGrid and MultipleShootingGrid (ShootingGridBase::findNearestShootingInterval + warmStartShifting,
 * shooting_grid_base.cpp:292-381: the same rule on the shooting nodes) structures.  NonUniformFiniteDifferencesVariableGrid returns
 * B200SQP_ERR_UNSUPPORTED: the reference never shifts it (isMovingHorizonWarmStartActive() is false,
 * non_uniform_finite_differences_variable_grid.h:79).  Follow with b200sqp_solve / b200sqp_step(cold_start=0). */
int b200sqp_warm_start_shift(b200sqp_handle h, const double* x0_new, int32_t* num_shift);

/* ---- the hot path ------------------------------------------------------------------------------------------------------ */
/* LevenbergMarquardtSparse::solve (levenberg_marquardt_sparse.cpp:44-220) for all instances, entirely on the device:
 * opts->iterations outer passes each; new_run=1 resets the penalty weights, 0 adapts them (:83-86).
 * status [batch] (b200sqp_status) and chi2 [batch] (*obj_value) are host pointers and may be NULL. */
int b200sqp_solve(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t new_run, int32_t* status, double* chi2);
/* Same, but nothing crosses PCIe: results stay in HBM (b200sqp_get_* fetches them).  Asynchronous on the handle's stream. */
int b200sqp_solve_async(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t new_run);
int b200sqp_synchronize(b200sqp_handle h);
/* One call a controller makes per MPC step for a whole batch: H2D of x0 [batch*nx] (+xref), trajectory initialisation when
 * `cold_start`, solve, D2H of the optimised parameter vectors [batch*n_params], chi2 and status.  Host pointers (pinned or not). */
int b200sqp_step(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t cold_start, const double* x0, const double* xref,
                 double* params_out, double* chi2_out, int32_t* status_out);

/* One closed-loop MPC step of the whole batch with minimal host traffic: what PredictiveController::step
 * (src/controllers/src/predictive_controller.cpp:46-79) needs from StructuredOptimalControlProblem::compute + getFirstControlInput.
 * H2D of the measured states x0 [batch*nx] (+ xref or NULL), then per `mode`
 *   0  cold start: FullDiscretizationGridBase::initializeSequences
 *   1  keep the previous solution as the initial guess, start state replaced (grid update with warm start off)
 *   2  moving-horizon warm start: b200sqp_warm_start_shift (fixed-dt grids)
 * solve (new_run = 1), D2H of the first controls u0_out [batch*nu] and, if non-NULL, chi2_out [batch] / status_out [batch].
 * The trajectories stay in HBM (b200sqp_get_params fetches them). */
int b200sqp_mpc_step(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t mode, const double* x0, const double* xref, double* u0_out,
                     double* chi2_out, int32_t* status_out);

/* Plant simulation for `batch` independent plants: SimulatedPlant::control
 * (src/plants/src/simulated_plant.cpp:92-146; no dead time, no disturbances) = one solveIVP of the system dynamics over `dt` with the
 * control held, integrator 0 = IntegratorExplicitEuler (the plant's default, simulated_plant.cpp:37;
 * src/numerics/include/corbo-numerics/explicit_integrators.h:66-72), 1 = IntegratorExplicitRungeKutta4 (explicit_integrators.h:280-295);
 * same expression order as the reference, no FMA contraction.  Host pointers: x [batch*nx], u [batch*nu], x_next [batch*nx].
 * Handle-less like b200sqp_linearize_dynamics. */
int b200sqp_plant_step(int32_t dynamics, const double* dyn_params, int32_t integrator, double dt, int32_t batch, const double* x, const double* u,
                       double* x_next, int32_t device);

/* The whole closed loop of the batch on the device -- what ClosedLoopControlTask::performTask
 * (src/tasks/src/task_closed_loop_control.cpp:153-235) does per instance with a PredictiveController and a SimulatedPlant, and what
 * BenchmarkTaskVaryingInitialState repeats over start states: for s = 0..steps-1
 *   measurement = plant state (full-state output, no observer dynamics);
 *   controller  = b200sqp_mpc_step's device half (step 0 always initialises the grid; later steps per `mode`: 0 re-initialise,
 *                 1 keep the previous solution, 2 moving-horizon warm start), first control of the optimised trajectory;
 *   plant       = b200sqp_plant_step's kernel over `plant_dt` with the dynamics and parameters of the handle's OCP.
 * Only x0 [batch*nx] (+ xref [batch*nx] or NULL) go in; after the last step the log comes out (each may be NULL):
 * u_applied [steps][batch*nu], x_closed [steps+1][batch*nx] (row 0 = x0), chi2_out / status_out [steps][batch].
 * Nothing crosses PCIe between the steps. */
int b200sqp_closed_loop(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t mode, int32_t integrator, double plant_dt, int32_t steps,
                        const double* x0, const double* xref, double* u_applied, double* x_closed, double* chi2_out, int32_t* status_out);

/* LevenbergMarquardtSparse::computeValues (:222-246) and ...EdgeBased::computeCombinedSparseJacobian (:1480-1753) at the current
 * parameters, penalty weights applied: values [batch*m], jac_values [batch*nnzJ] in the CSC order of b200sqp_jacobian_pattern.
 * Like the reference, evaluating the Jacobian perturbs the parameters in place (+d,-2d,+d; edge_interface.cpp:78-85). */
int b200sqp_evaluate(b200sqp_handle h, double weight_eq, double weight_ineq, double weight_bounds, double* values, double* jac_values);

/* Numerical linearisation of the system dynamics at `batch` points, A_i = df/dx, B_i = df/du at (x_i, u_i):
 * SystemDynamicsInterface::getLinearA / getLinearB (src/systems/src/system_dynamics_interface.cpp:33-59) with the reference's
 * finite-difference rules, method 0 = ForwardDifferences (the reference's default, finite_differences.hpp:29-48), 1 =
 * CentralDifferences (finite_differences.hpp:167-188); delta = 1e-9, in-place perturbation, no FMA contraction.
 * Host pointers: x [batch*nx], u [batch*nu]; A [batch*nx*nx], B [batch*nx*nu] column-major per point, either may be NULL.
 * Handle-less: `dynamics` is a b200sqp_dynamics id, dyn_params its parameters [B200SQP_MAX_DYN_PARAMS]. */
int b200sqp_linearize_dynamics(int32_t dynamics, const double* dyn_params, int32_t method, int32_t batch, const double* x, const double* u,
                               double* A, double* B, int32_t device);

/* Finite-difference Hessian of the system dynamics w.r.t. z = [x; u] at `batch` points, the remaining half of the reference's
 * numerics/finite_differences: method 0 = ForwardDifferences::hessian
 * (src/numerics/include/corbo-numerics/finite_differences.hpp:50-104), 1 = CentralDifferences::hessian (:190-273); delta = 1e-5,
 * in-place increments in the reference's order, every (i, j) pair evaluated.  H_p = sum_v multipliers[p][v] * d2 f_v / dz dz
 * (multipliers [batch*nx], NULL = plain sum over the components, as in the reference); H [batch*(nx+nu)^2], column-major per point.
 * Host pointers; handle-less like b200sqp_linearize_dynamics. */
int b200sqp_dynamics_hessian(int32_t dynamics, const double* dyn_params, int32_t method, int32_t batch, const double* x, const double* u,
                             const double* multipliers, double* H, int32_t device);

/* Per-instance LM bookkeeping of the last solve: [batch] each, host pointers, any may be NULL.
 * inner_passes = number of factorisations, rejects = rejected trial steps, relinearizations = Jacobian evaluations. */
int b200sqp_get_statistics(b200sqp_handle h, int32_t* inner_passes, int32_t* rejects, int32_t* relinearizations, double* mu, double* rho);
/* per-iteration chi2 trace of the last solve, [batch*(iterations+1)] (row = instance); entry 0 is the initial chi2 */
int b200sqp_get_chi2_trace(b200sqp_handle h, double* trace, int32_t iterations);

/* ---- measurement / multi-GPU plumbing ---------------------------------------------------------------------------------- */
/* device time of the last solve (CUDA events on the handle's stream), milliseconds */
int b200sqp_last_solve_ms(b200sqp_handle h, float* ms);
/* number of kernel launches issued by this handle so far */
int b200sqp_launch_count(b200sqp_handle h, int64_t* launches);
/* raw device pointers for zero-copy interop (torch / NCCL all-gather of the stop-test residuals): chi2 [batch] doubles,
 * status [batch] int32, x0 [batch*nx] doubles.  Valid until destroy. */
int b200sqp_device_pointers(b200sqp_handle h, void** chi2, void** status, void** x0);
/* Fused stop-test exchange over NVLink peer memory (one process per GPU on one NVSwitch box).  Replaces the
 * separate all-gather launch after each solve: once attached, the LM kernel of every rank stores its per-instance chi2 directly
 * into every rank's gather buffer (peer stores) and signals arrival; b200sqp_peer_wait (stream-ordered, bounded spin) returns once
 * all ranks' values of the last solve are present in this rank's buffer.
 *   1. every rank: b200sqp_peer_export -> 64-byte CUDA IPC handle of its gather buffer
 *   2. exchange the handles (e.g. torch.distributed all_gather), every rank: b200sqp_peer_attach(all handles in rank order)
 *   3. per step: b200sqp_solve_async; b200sqp_peer_wait; read gathered[world*batch] (global instance order) via b200sqp_peer_gathered
 * Equal `batch` on all ranks.  world <= 8. */
#define B200SQP_IPC_HANDLE_BYTES 64
int b200sqp_peer_export(b200sqp_handle h, int32_t world, int32_t rank, void* ipc_handle_out /*[64]*/);
/* B200SQP_ERR_INVALID when already attached (detach first) */
int b200sqp_peer_attach(b200sqp_handle h, const void* ipc_handles /*[world*64], rank order; own entry ignored*/);
/* enqueue the bounded wait for the gathered chi2 of the last solve; returns B200SQP_ERR_CUDA from a later call if it timed out */
int b200sqp_peer_wait(b200sqp_handle h);
/* device pointer to this rank's gathered chi2 of the last solve, [world*batch] doubles */
int b200sqp_peer_gathered(b200sqp_handle h, void** chi2_all);
/* synchronises the stream and reports whether any b200sqp_peer_wait since attach hit its 2 s bound (a rank that never arrived) */
int b200sqp_peer_status(b200sqp_handle h, int32_t* timed_out);
int b200sqp_peer_detach(b200sqp_handle h);
/* tuning knob: cooperating threads per instance in the LM kernel (1, 2, 4, 8; 0 = choose from the batch size).  Structures with
 * large stage blocks (quadrotor) run a warp-per-instance pipeline instead of the fused kernel; -1 forces the fused kernel there. */
int b200sqp_set_threads_per_instance(b200sqp_handle h, int32_t threads);
/* test knob: the LM kernel is compiled per structure family in two feature sets -- the general one (state bounds, partially fixed goal,
 * final-stage constraint, any stage cost) and a lean one for structures that need none of these (picked automatically).  general = 1
 * forces the general set on a lean-eligible structure so that parity tests can run both variants on the same problem; 0 = automatic. */
int b200sqp_set_feature_set(b200sqp_handle h, int32_t general);
/* measurement aid: when enabled, thread 0 of every thread block of the LM kernel accumulates clock64() per phase; get returns the
 * mean over thread blocks of the last solve, mean_cycles[5] = {linearise (a3/a4/a13), factor+solve (a14), trial values (a2/a12),
 * LM control (a1), stop-test exchange epilogue (peer stores + fence + arrivals; 0 without attached peers)} in SM clock cycles.  Off by
 * default (the kernel then only tests one pointer). */
int b200sqp_set_phase_profile(b200sqp_handle h, int32_t enable);
int b200sqp_get_phase_cycles(b200sqp_handle h, double* mean_cycles /*[5]*/);
/* Arithmetic of the solve.  The reference is fp64 only and so is every path here by default.  B200SQP_PRECISION_F32 is the reduced-
 * precision variant BASELINE.json configs[4] names (12-state quadrotor): Jacobian columns by central differences with delta = 2^-10 in
 * fp32, normal equations, Cholesky factor and substitutions in fp32; parameters, steps, trial-point residuals and the LM control state
 * stay fp64.  It has no reference counterpart: parity is judged against the fp64 path at 1e-3 relative on the trajectories (SURVEY.md
 * section 8d).  Available for structures that run the warp-cooperative pipeline, else B200SQP_ERR_UNSUPPORTED. */
typedef enum { B200SQP_PRECISION_F64 = 0, B200SQP_PRECISION_F32 = 1 } b200sqp_precision;
int b200sqp_set_precision(b200sqp_handle h, int32_t precision);
/* measurement aid: fp64 FMA throughput of `device` in TFLOP/s (2 flops per FMA), measured live with a register-resident kernel of
 * independent FMA chains on every SM -- the issue bound the fused LM kernel is reported against next to the HBM roofline
 * (SURVEY.md section 8d "also report the fp64 FMA bound"). */
int b200sqp_measure_fp64_peak(int32_t device, double* tflops);
/* make the handle launch on an external stream (e.g. torch's current stream); pass NULL to restore its own */
int b200sqp_set_stream(b200sqp_handle h, void* cuda_stream);

/* ---- time-optimal grids with grid adaptation: per-instance grid size ---------------------------------------
 * Replaces, for `batch` independent controllers, NonUniformFiniteDifferencesVariableGrid with setGridAdaptTimeBasedSingleStep(n_max,
 * dt_hyst_ratio) + setNmin(n_min) + setWarmStart(warm_start)
 * (src/optimal_control/include/corbo-optimal-control/structured_ocp/discretization_grids/non_uniform_finite_differences_variable_grid.h:52-54,
 * src/optimal_control/src/structured_ocp/discretization_grids/non_uniform_finite_differences_variable_grid.cpp:176-257) under the OCP loop of
 * PredictiveController::step (src/controllers/src/predictive_controller.cpp:66).  Instances are bucketed by grid size on the device; a
 * bucket is an ordinary solver of `ocp` with n_grid = N, created on first use and sized for the whole batch (180 GB of HBM make that the
 * cheap choice).  ocp->grid must be B200SQP_GRID_FD_NONUNIFORM_VARDT; ocp->n_grid is the initial size of every instance;
 * 3 <= n_min <= n_max (the reference's default n_min = 2 leaves a single interval between two fixed states). */
typedef struct b200sqp_adaptive* b200sqp_adaptive_handle;
int b200sqp_adaptive_create(const b200sqp_ocp* ocp, int32_t batch, int32_t device, int32_t n_min, int32_t n_max, double dt_hyst_ratio,
                            int32_t warm_start, b200sqp_adaptive_handle* out);
int b200sqp_adaptive_destroy(b200sqp_adaptive_handle a);
/* One controller step of every instance = num_ocp_iterations x StructuredOptimalControlProblem::compute: the first with new_run = 1 (start
 * state <- x0, fixed goal components <- xref, penalty weights reset, no adaptation), the others with new_run = 0 (adaptGridTimeBasedSingleStep
 * on the previous solution -- one inserted or removed grid point per instance -- then the solve with adapted weights).  With warm_start = 0
 * every solve starts from initializeSequences at the instance's current grid size, as the reference does.  The very first solve of a handle
 * always initialises.  Host pointers: x0, xref [batch*nx]; u0_out [batch*nu], chi2_out / status_out / n_out [batch] (any may be NULL). */
int b200sqp_adaptive_step(b200sqp_adaptive_handle a, const b200sqp_lm_options* opts, int32_t num_ocp_iterations, const double* x0,
                          const double* xref, double* u0_out, double* chi2_out, int32_t* status_out, int32_t* n_out);
/* switch the handle to the reference's second strategy, setGridAdaptRedundantControls(n_max, num_backup_nodes, epsilon)
 * (non_uniform_finite_differences_variable_grid.cpp:52-58, 259-352): intervals whose control repeats in the successor (within epsilon) or
 * whose dt is below 1e-6 are redundant; a surplus over num_backup_nodes is removed from the back, a deficit is made up by halving the
 * interval with the largest dt -- several grid points per call, replayed per instance as an edit script on the device.  Call before the
 * first step; n_max <= 129.  (dt_hyst_ratio of b200sqp_adaptive_create is then unused.) */
int b200sqp_adaptive_set_redundant_controls(b200sqp_adaptive_handle a, int32_t num_backup_nodes, double epsilon);
/* count [batch]: how many adaptations of each instance so far changed its LAST interval.  The reference has no defined answer there (it
 * indexes one past the end of its vertex vectors: non_uniform_finite_differences_variable_grid.cpp:225 `_x_seq[i + 1]`, :237
 * `_dt_seq[i + 1]` with i = size - 1 -- stale memory at best, heap corruption at worst); the device uses x_f as the right neighbour of a
 * split last interval and drops the dt of a removed one.  Parity with the reference is claimed for instances whose count is 0. */
int b200sqp_adaptive_last_interval_changes(b200sqp_adaptive_handle a, int32_t* count);
/* create the buckets of grid sizes n_from..n_to (clipped to the reachable range) now instead of on first use: a bucket allocates its
 * device state for the whole batch, which a latency-sensitive control loop wants outside its first steps */
int b200sqp_adaptive_reserve(b200sqp_adaptive_handle a, int32_t n_from, int32_t n_to);
/* getStateAndControlTimeSeries of every instance, padded to n_cap grid points: x [batch][n_cap][nx] (N rows used), u [batch][n_cap][nu] and
 * dt [batch][n_cap] (N-1 rows used), n [batch]; unused rows are zero.  n_cap must cover the largest grid of the batch. */
int b200sqp_adaptive_get_trajectories(b200sqp_adaptive_handle a, int32_t n_cap, double* x, double* u, double* dt, int32_t* n);
/* bookkeeping: buckets in use, grid points inserted / removed so far, kernels launched */
int b200sqp_adaptive_statistics(b200sqp_adaptive_handle a, int32_t* occupied_buckets, int64_t* splits, int64_t* merges, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* B200SQP_H_ */
