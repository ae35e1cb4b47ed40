// Host-side structure of one OCP: dimensions, the reference's vertex/edge indexing, the CSC pattern of the combined Jacobian
// and the device-side block layout.  No CUDA in here: these functions back the handle-less part of the C ABI
// (b200sqp_dims_of / _vertex_indices / _edge_indices / _jacobian_pattern) and run on machines without a GPU.
//
// Indexing rules restated from the reference (paths relative to /root/reference/src):
//   vertex index = prefix sum of unfixed dimensions over the active vertices       optimization/src/hyper_graph/vertex_set.cpp:405-418
//   active order = [x_k?, u_k, dt_k?] per interval, then xf                         optimal_control/src/structured_ocp/discretization_grids/
//                                                                                   full_discretization_grid_base.cpp:514-527,
//                                                                                   non_uniform_full_discretization_grid_base.cpp:454-467,
//                                                                                   shooting_grid_base.cpp:583-598
//   edge index   = prefix sum of edge dimensions per category in creation order     optimization/src/hyper_graph/edge_set.cpp:31-42,101-166
//   creation order per interval: state cost, control cost, dt cost (twice), dynamics  optimal_control/src/functions/nlp_functions.cpp:70-132,
//                                                                                   .../finite_differences_grid.cpp:38-154
//   Jacobian rows: lsq | equalities | inequalities | bounds                         optimization/src/hyper_graph/
//                                                                                   hyper_graph_optimization_problem_edge_based.cpp:1491-1493
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200sqp.h"

namespace b200sqp {

constexpr double kCorboInf = 2e30;  // CORBO_INF_DBL, core/include/corbo-core/types.h:53

// Square root of a cost weight the way the reference takes it (QuadraticFormCost::setWeightQ / setWeightR, quadratic_cost.cpp:32-96;
// QuadraticFinalStateCost::setWeightQf, final_state_cost.cpp:38-69): diagonal to 1e-10 -> element-wise square root of the diagonal;
// else the UPPER Cholesky factor U, M = U^T U (Eigen::LLT<MatrixXd, Upper>, unblocked below 32 rows).
struct WeightSqrt
{
    bool dense = false;
    std::vector<double> w;  // diagonal: [dim]; dense: [dim*dim] row-major, zeros below the diagonal
    bool ok = true;          // false: not positive definite (LLT reports NumericalIssue -> the reference's setter returns false)
};
WeightSqrt weightSqrt(const double* diag, const double* full, int dense_flag, int dim);

struct Structure
{
    b200sqp_ocp ocp;
    int K = 0;         // intervals = N-1
    int nx = 0, nu = 0;
    int vt = 0;        // 1: one free dt per interval (non-uniform grid)
    int nb = 0;        // device block dimension nu + vt + nx, block k = [u_k, dt_k, x_{k+1}]
    int defect = 0;    // DEFECT_* id (dynamics.cuh)
    b200sqp_dims dims{};

    // reference indexing
    std::vector<int32_t> x_idx, u_idx, dt_idx;                 // [N], [K], [K]; -1 = fixed
    std::vector<int32_t> state_cost_idx, control_cost_idx;     // [K] lsq row offsets, -1 = absent
    std::vector<int32_t> dt_cost_idx;                          // [2K]
    std::vector<int32_t> dynamics_idx;                         // [K] equality row offsets
    std::vector<int32_t> dt_eq_idx;                            // [K] equality row offset of TwoScalarEqualEdge(dt_{k-1}, dt_k), -1 if none
    int dteq = 0;                                              // 1: consecutive dt vertices are coupled by equality edges
    int32_t final_cost_idx = -1;
    int32_t final_eq_idx = -1, final_ineq_idx = -1;            // final-stage constraint edge: row offset inside its category
    std::vector<int32_t> bound_row;                            // [n] row offset inside the bounds block or -1
    std::vector<int32_t> col_ptr, row_idx;                     // CSC pattern of the combined Jacobian

    // device layout <-> reference parameter order
    std::vector<int32_t> ref_of_internal;                      // [K*nb] reference parameter index of internal slot, -1 = fixed component
    std::vector<int32_t> internal_of_ref;                      // [n]

    // per-interval scatter tables for b200sqp_evaluate (values rows and CSC positions), see lm_device.cuh
    WeightSqrt q_w, r_w, qf_w;                                  // square roots of the stage / control / final weights
    bool denseCost() const { return q_w.dense || r_w.dense || qf_w.dense; }
    std::vector<int32_t> value_rows;                           // [K * values_per_interval]
    std::vector<int32_t> jac_pos;                              // [K * jac_per_interval]
    int values_per_interval = 0, jac_per_interval = 0;

    bool xfFixed(int i) const { return ocp.xf_fixed[i] != 0; }
    bool xfFullyFixed() const
    {
        for (int i = 0; i < nx; ++i)
            if (!xfFixed(i)) return false;
        return true;
    }
};

// (nx, nu) of a b200sqp_dynamics id; false if the id is not in the registry
bool dynamicsDimensions(int dynamics, int& nx, int& nu);

// Validates the descriptor against the closed registry and fills everything above.  Returns B200SQP_OK or an error code and
// leaves a message in `err`.
int buildStructure(const b200sqp_ocp& ocp, Structure& s, std::string& err);

// layout of the per-interval tables (shared by structure.cpp and the materialising kernel)
struct EvalLayout
{
    int nx, nu, vt;
    // values rows per interval: [x0 cost nx | control cost nu | dt cost 2 | cost on x_{k+1} nx | defect nx | bounds u nu | bound dt 1 |
    //                            bounds x_{k+1} nx | final-stage equality nx | final-stage inequality 1]
    int v_x0c() const { return 0; }
    int v_uc() const { return nx; }
    int v_tc() const { return nx + nu; }
    int v_xs() const { return nx + nu + 2; }
    int v_e() const { return 2 * nx + nu + 2; }
    int v_ub() const { return 3 * nx + nu + 2; }
    int v_tb() const { return 3 * nx + 2 * nu + 2; }
    int v_xb() const { return 3 * nx + 2 * nu + 3; }
    int v_teq() const { return 4 * nx + 2 * nu + 3; }  // final-stage equality rows (last interval only)
    int v_tin() const { return 5 * nx + 2 * nu + 3; }  // final-stage inequality row
    int v_dq() const { return 5 * nx + 2 * nu + 4; }   // TwoScalarEqualEdge(dt_{k-1}, dt_k) row
    int v_count() const { return 5 * nx + 2 * nu + 5; }
    // Jacobian positions per interval: [uc diag nu | tc 2 | xs diag nx | A nx*nx (col-major) | Bu nx*nu | Bt nx | C nx*nx |
    //                                   bounds u nu | bound dt 1 | bounds x_{k+1} nx | final-stage equality nx*nx | final-stage inequality nx]
    int j_uc() const { return 0; }
    int j_tc() const { return nu; }
    int j_xs() const { return nu + 2; }
    int j_A() const { return nu + 2 + nx; }
    int j_Bu() const { return j_A() + nx * nx; }
    int j_Bt() const { return j_Bu() + nx * nu; }
    int j_C() const { return j_Bt() + nx; }
    int j_ub() const { return j_C() + nx * nx; }
    int j_tb() const { return j_ub() + nu; }
    int j_xb() const { return j_tb() + 1; }
    int j_teq() const { return j_xb() + nx; }        // nx*nx (col-major), final-stage equality block
    int j_tin() const { return j_teq() + nx * nx; }  // nx, final-stage inequality row
    // full (non-diagonal) cost weights: the whole nu x nu / nx x nx block of the control-cost edge and of the cost edge on x_{k+1}
    // (col-major: [column c][row i]); the diagonal slots above stay unused then
    int j_ucd() const { return j_tin() + nx; }
    int j_xsd() const { return j_ucd() + nu * nu; }
    // TwoScalarEqualEdge(dt_{k-1}, dt_k), stored with interval k: d/d dt_{k-1}, d/d dt_k
    int j_dq() const { return j_xsd() + nx * nx; }
    int j_count() const { return j_dq() + 2; }
};



}  // namespace b200sqp
