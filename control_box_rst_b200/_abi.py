"""ctypes mirror of include/b200sqp.h (the C ABI of libb200sqp.so).

Only plain structs and enums live here so that the test-side oracle bindings (tests/oracle_bindings.py) can share the
descriptor types without importing any product code path.
"""
import ctypes as C

MAX_NX = 16
MAX_NU = 8
MAX_DYN_PARAMS = 32

CORBO_INF_DBL = 2e30  # core/include/corbo-core/types.h:53

# b200sqp_dynamics
DYN_VAN_DER_POL = 0
DYN_DUFFING = 1
DYN_SIMPLE_PENDULUM = 2
DYN_CART_POLE = 3
DYN_DOUBLE_INTEGRATOR = 4
DYN_UNICYCLE = 5
DYN_QUADROTOR = 6
DYN_FREE_SPACE_ROCKET = 7
DYN_MASSLESS_PENDULUM = 8
DYN_TOY_EXAMPLE = 9
DYN_ARTSTEINS_CIRCLE = 10
DYN_LINEAR_2X1 = 11
DYN_LINEAR_3X1 = 12
DYN_LINEAR_4X1 = 13
DYN_LINEAR_4X2 = 14
DYN_TRIPLE_INTEGRATOR = 15
DYN_QUAD_INTEGRATOR = 16
DYN_DIMS = {  # id -> (nx, nu)
    DYN_VAN_DER_POL: (2, 1),
    DYN_DUFFING: (2, 1),
    DYN_SIMPLE_PENDULUM: (2, 1),
    DYN_CART_POLE: (4, 1),
    DYN_DOUBLE_INTEGRATOR: (2, 1),
    DYN_UNICYCLE: (3, 2),
    DYN_QUADROTOR: (12, 4),
    DYN_FREE_SPACE_ROCKET: (3, 1),
    DYN_MASSLESS_PENDULUM: (2, 1),
    DYN_TOY_EXAMPLE: (2, 1),
    DYN_ARTSTEINS_CIRCLE: (2, 1),
    DYN_LINEAR_2X1: (2, 1),
    DYN_LINEAR_3X1: (3, 1),
    DYN_LINEAR_4X1: (4, 1),
    DYN_LINEAR_4X2: (4, 2),
    DYN_TRIPLE_INTEGRATOR: (3, 1),
    DYN_QUAD_INTEGRATOR: (4, 1),
}

# b200sqp_grid
GRID_FD_UNIFORM = 0
GRID_FD_NONUNIFORM_VARDT = 1
GRID_MULTIPLE_SHOOTING = 2

# b200sqp_collocation
COLL_FORWARD = 0
COLL_BACKWARD = 1
COLL_MIDPOINT = 2
COLL_CRANK_NICOLSON = 3

# b200sqp_integrator
INT_EULER = 0
INT_RK4 = 1

# b200sqp_stage_cost
COST_NONE = 0
COST_QUADRATIC_LSQ = 1
COST_MINIMUM_TIME_LSQ = 2

# b200sqp_final_constraint
FINAL_CONSTRAINT_NONE = 0
FINAL_CONSTRAINT_EQUALITY = 1
FINAL_CONSTRAINT_BALL = 2

# b200sqp_status == corbo::SolverStatus
STATUS_CONVERGED = 0
STATUS_EARLY_TERMINATED = 1
STATUS_INFEASIBLE = 2
STATUS_ERROR = 3

ERR_INVALID = -1
ERR_UNSUPPORTED = -2
ERR_CUDA = -3
ERR_NO_DEVICE = -4
ERR_NOT_LSQ = -5


class Ocp(C.Structure):
    """struct b200sqp_ocp"""

    _fields_ = [
        ("grid", C.c_int32),
        ("dynamics", C.c_int32),
        ("collocation", C.c_int32),
        ("integrator", C.c_int32),
        ("n_grid", C.c_int32),
        ("nx", C.c_int32),
        ("nu", C.c_int32),
        ("stage_cost", C.c_int32),
        ("final_cost", C.c_int32),
        ("zero_x_ref", C.c_int32),
        ("zero_u_ref", C.c_int32),
        ("xf_fixed", C.c_int32 * MAX_NX),
        ("dt_ref", C.c_double),
        ("dt_lb", C.c_double),
        ("dt_ub", C.c_double),
        ("dyn_params", C.c_double * MAX_DYN_PARAMS),
        ("q_diag", C.c_double * MAX_NX),
        ("r_diag", C.c_double * MAX_NU),
        ("qf_diag", C.c_double * MAX_NX),
        ("x_lb", C.c_double * MAX_NX),
        ("x_ub", C.c_double * MAX_NX),
        ("u_lb", C.c_double * MAX_NU),
        ("u_ub", C.c_double * MAX_NU),
        ("final_constraint", C.c_int32),
        ("term_xref", C.c_double * MAX_NX),
        ("term_s_diag", C.c_double * MAX_NX),
        ("term_gamma", C.c_double),
        ("q_dense", C.c_int32),
        ("r_dense", C.c_int32),
        ("qf_dense", C.c_int32),
        ("q_full", C.c_double * (MAX_NX * MAX_NX)),
        ("r_full", C.c_double * (MAX_NU * MAX_NU)),
        ("qf_full", C.c_double * (MAX_NX * MAX_NX)),
        ("dt_eq_constraint", C.c_int32),
    ]


class LmOptions(C.Structure):
    """struct b200sqp_lm_options; defaults of LevenbergMarquardtSparse (levenberg_marquardt_sparse.h:112-124)"""

    _fields_ = [
        ("iterations", C.c_int32),
        ("weight_eq", C.c_double),
        ("weight_ineq", C.c_double),
        ("weight_bounds", C.c_double),
        ("adapt_factor_eq", C.c_double),
        ("adapt_factor_ineq", C.c_double),
        ("adapt_factor_bounds", C.c_double),
        ("adapt_max_eq", C.c_double),
        ("adapt_max_ineq", C.c_double),
        ("adapt_max_bounds", C.c_double),
    ]

    @classmethod
    def defaults(cls, iterations=10, weights=(2.0, 2.0, 2.0), factors=(1.0, 1.0, 1.0), maxima=(500.0, 500.0, 500.0)):
        return cls(iterations, *weights, *factors, *maxima)


class Dims(C.Structure):
    """struct b200sqp_dims"""

    _fields_ = [
        ("n_params", C.c_int32),
        ("m_lsq", C.c_int32),
        ("m_eq", C.c_int32),
        ("m_ineq", C.c_int32),
        ("m_bounds", C.c_int32),
        ("nnz_jacobian", C.c_int32),
        ("nnz_hessian_upper", C.c_int32),
        ("n_blocks", C.c_int32),
        ("block_dim", C.c_int32),
        ("algorithmic_bytes_per_iteration", C.c_int64),
    ]

    @property
    def m(self):
        return self.m_lsq + self.m_eq + self.m_ineq + self.m_bounds
