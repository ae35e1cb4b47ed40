// Host-side structure derivation (no CUDA).  See structure.h for the reference rules that are restated here.
#include "structure.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "dynamics_ids.h"

namespace b200sqp {

namespace {

bool finiteLb(double lb) { return lb > -kCorboInf; }  // vector_vertex.h:174-178
bool finiteUb(double ub) { return ub < kCorboInf; }   // vector_vertex.h:180-184

struct Dim
{
    int nx, nu;
};

bool dynamicsDims(int id, Dim& d)
{
    switch (id)
    {
        case B200SQP_DYN_VAN_DER_POL:
        case B200SQP_DYN_DUFFING:
        case B200SQP_DYN_SIMPLE_PENDULUM:
        case B200SQP_DYN_DOUBLE_INTEGRATOR:
        case B200SQP_DYN_MASSLESS_PENDULUM:
        case B200SQP_DYN_TOY_EXAMPLE:
        case B200SQP_DYN_ARTSTEINS_CIRCLE:
        case B200SQP_DYN_LINEAR_2X1:
            d = {2, 1};
            return true;
        case B200SQP_DYN_FREE_SPACE_ROCKET:
        case B200SQP_DYN_LINEAR_3X1:
        case B200SQP_DYN_TRIPLE_INTEGRATOR:
            d = {3, 1};
            return true;
        case B200SQP_DYN_LINEAR_4X1:
        case B200SQP_DYN_QUAD_INTEGRATOR:
            d = {4, 1};
            return true;
        case B200SQP_DYN_LINEAR_4X2:
            d = {4, 2};
            return true;
        case B200SQP_DYN_CART_POLE:
            d = {4, 1};
            return true;
        case B200SQP_DYN_UNICYCLE:
            d = {3, 2};
            return true;
        case B200SQP_DYN_QUADROTOR:
            d = {12, 4};
            return true;
    }
    return false;
}

}  // namespace

bool dynamicsDimensions(int dynamics, int& nx, int& nu)
{
    Dim d;
    if (!dynamicsDims(dynamics, d)) return false;
    nx = d.nx;
    nu = d.nu;
    return true;
}

WeightSqrt weightSqrt(const double* diag, const double* full, int dense_flag, int dim)
{
    WeightSqrt r;
    auto fromDiagonal = [&](const double* d, int stride) {
        r.w.resize(dim);
        for (int i = 0; i < dim; ++i) r.w[i] = std::sqrt(d[(size_t)i * stride]);
    };
    if (!dense_flag)
    {
        fromDiagonal(diag, 1);
        return r;
    }
    double max_diag = 0;
    for (int i = 0; i < dim; ++i) max_diag = std::max(max_diag, std::fabs(full[i * dim + i]));
    bool is_diag = true, is_zero = true;
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
        {
            if (i != j && !(std::fabs(full[i * dim + j]) <= max_diag * 1e-10)) is_diag = false;  // MatrixBase::isDiagonal(1e-10)
            if (!(std::fabs(full[i * dim + j]) <= 1e-12)) is_zero = false;                      // MatrixBase::isZero()
        }
    if (is_diag)
    {
        fromDiagonal(full, dim + 1);
        for (double v : r.w) r.ok = r.ok && !(v != v);
        return r;
    }
    r.dense = true;
    r.w.assign((size_t)dim * dim, 0.0);
    if (is_zero) return r;
    // llt_inplace<double, Lower>::unblocked on the transposed view (LLT.h:290-320): column k of L = U^T from the upper triangle of M
    std::vector<double> L((size_t)dim * dim, 0.0);
    for (int k = 0; k < dim && r.ok; ++k)
    {
        double x = full[k * dim + k];
        if (k > 0)
        {
            double sq = L[k * dim] * L[k * dim];
            for (int j = 1; j < k; ++j) sq += L[k * dim + j] * L[k * dim + j];
            x -= sq;
        }
        if (!(x > 0))
        {
            r.ok = false;
            break;
        }
        x              = std::sqrt(x);
        L[k * dim + k] = x;
        for (int i = k + 1; i < dim; ++i)
        {
            double a = full[k * dim + i];
            if (k > 0)
            {
                double dot = L[i * dim] * L[k * dim];
                for (int j = 1; j < k; ++j) dot += L[i * dim + j] * L[k * dim + j];
                a -= dot;
            }
            L[i * dim + k] = a / x;
        }
    }
    for (int i = 0; i < dim; ++i)
        for (int j = i; j < dim; ++j) r.w[i * dim + j] = L[j * dim + i];
    return r;
}

int buildStructure(const b200sqp_ocp& ocp, Structure& s, std::string& err)
{
    s     = Structure();
    s.ocp = ocp;
    Dim dd;
    if (!dynamicsDims(ocp.dynamics, dd))
    {
        err = "unknown dynamics id (closed functor registry; no CPU fallback)";
        return B200SQP_ERR_UNSUPPORTED;
    }
    if (ocp.nx != dd.nx || ocp.nu != dd.nu)
    {
        err = "nx/nu do not match the dynamics id";
        return B200SQP_ERR_INVALID;
    }
    if (ocp.dynamics == B200SQP_DYN_CART_POLE)
    {
        // CartPoleSystem has no parameter setters (nonlinear_benchmark_systems.h:337-365: private constants): the device functor uses the
        // same constants.  All-zero parameters mean "the class as it is"; anything else must be those constants.
        const double fixed[4] = {1.0, 0.3, 0.5, 9.81};
        bool zero = true, same = true;
        for (int i = 0; i < 4; ++i)
        {
            zero = zero && ocp.dyn_params[i] == 0.0;
            same = same && ocp.dyn_params[i] == fixed[i];
        }
        if (!zero && !same)
        {
            err = "the cart-pole of the reference has fixed parameters (mc, mp, l, g) = (1, 0.3, 0.5, 9.81); other values are not supported";
            return B200SQP_ERR_UNSUPPORTED;
        }
    }
    if (ocp.n_grid < 2)
    {
        err = "n_grid must be >= 2";
        return B200SQP_ERR_INVALID;
    }
    if (!(ocp.dt_ref > 0))
    {
        err = "dt_ref must be > 0";
        return B200SQP_ERR_INVALID;
    }
    switch (ocp.grid)
    {
        case B200SQP_GRID_FD_UNIFORM:
        case B200SQP_GRID_FD_NONUNIFORM_VARDT:
            if (ocp.collocation < B200SQP_COLL_FORWARD || ocp.collocation > B200SQP_COLL_CRANK_NICOLSON)
            {
                err = "unknown collocation id";
                return B200SQP_ERR_UNSUPPORTED;
            }
            s.defect = ocp.collocation;  // DEFECT_FORWARD..DEFECT_CRANK_NICOLSON share the collocation ids
            break;
        case B200SQP_GRID_MULTIPLE_SHOOTING:
            if (ocp.integrator != B200SQP_INT_EULER && ocp.integrator != B200SQP_INT_RK4)
            {
                err = "unknown integrator id";
                return B200SQP_ERR_UNSUPPORTED;
            }
            s.defect = ocp.integrator == B200SQP_INT_EULER ? DEFECT_EULER : DEFECT_RK4;
            break;
        default:
            err = "unknown grid id";
            return B200SQP_ERR_UNSUPPORTED;
    }
    if (ocp.stage_cost != B200SQP_COST_NONE && ocp.stage_cost != B200SQP_COST_QUADRATIC_LSQ && ocp.stage_cost != B200SQP_COST_MINIMUM_TIME_LSQ)
    {
        err = "unknown stage cost id";
        return B200SQP_ERR_UNSUPPORTED;
    }
    if (ocp.final_constraint < B200SQP_FINAL_CONSTRAINT_NONE || ocp.final_constraint > B200SQP_FINAL_CONSTRAINT_BALL)
    {
        err = "unknown final-stage constraint id";
        return B200SQP_ERR_UNSUPPORTED;
    }
    if (ocp.stage_cost == B200SQP_COST_QUADRATIC_LSQ && !ocp.zero_u_ref)
    {
        // quadratic_cost.cpp:161,163: the lsq branch with a non-zero control reference returns a scalar into a vector -> not a
        // least-squares form the reference itself handles consistently
        err = "QuadraticFormCost in lsq form requires a zero control reference";
        return B200SQP_ERR_NOT_LSQ;
    }
    if (ocp.stage_cost == B200SQP_COST_QUADRATIC_LSQ)
    {
        for (int i = 0; i < ocp.nx; ++i)
            if (!ocp.q_dense && ocp.q_diag[i] < 0)
            {
                err = "negative Q diagonal";
                return B200SQP_ERR_INVALID;
            }
        for (int i = 0; i < ocp.nu; ++i)
            if (!ocp.r_dense && ocp.r_diag[i] < 0)
            {
                err = "negative R diagonal";
                return B200SQP_ERR_INVALID;
            }
        s.q_w = weightSqrt(ocp.q_diag, ocp.q_full, ocp.q_dense, ocp.nx);
        s.r_w = weightSqrt(ocp.r_diag, ocp.r_full, ocp.r_dense, ocp.nu);
        if (!s.q_w.ok || !s.r_w.ok)
        {
            err = "Q / R is not positive definite (the reference's setWeightQ / setWeightR fail on it too)";
            return B200SQP_ERR_INVALID;
        }
        if (s.q_w.dense && ocp.zero_x_ref)
        {
            // quadratic_cost.cpp:112: lsq form + zero reference + non-diagonal Q assigns the SCALAR x^T U x to the cost vector
            err = "a non-diagonal Q needs a non-zero state reference: the reference's zero-reference lsq branch returns a scalar (quadratic_cost.cpp:112)";
            return B200SQP_ERR_UNSUPPORTED;
        }
    }
    if (ocp.final_cost == 1)
    {
        s.qf_w = weightSqrt(ocp.qf_diag, ocp.qf_full, ocp.qf_dense, ocp.nx);
        if (!s.qf_w.ok)
        {
            err = "Qf is not positive definite";
            return B200SQP_ERR_INVALID;
        }
    }

    const int nx = ocp.nx, nu = ocp.nu, N = ocp.n_grid, K = N - 1;
    s.nx = nx;
    s.nu = nu;
    s.K  = K;
    s.vt = ocp.grid == B200SQP_GRID_FD_NONUNIFORM_VARDT ? 1 : 0;
    s.nb = nu + s.vt + nx;
    const bool single_dt = !s.vt;  // hasSingleDt()
    if (ocp.dt_eq_constraint && !s.vt)
    {
        err = "dt_eq_constraint belongs to the NonUniformFiniteDifferencesVariableGrid";
        return B200SQP_ERR_INVALID;
    }
    s.dteq = (ocp.dt_eq_constraint && K > 1) ? 1 : 0;

    // ---- vertex indices ---------------------------------------------------------------------------------------------------
    s.x_idx.assign(N, -1);
    s.u_idx.assign(K, -1);
    s.dt_idx.assign(K, -1);
    int idx = 0;
    for (int k = 0; k < K; ++k)
    {
        if (k > 0)  // x_0 is fixed (full_discretization_grid_base.cpp:170)
        {
            s.x_idx[k] = idx;
            idx += nx;
        }
        s.u_idx[k] = idx;
        idx += nu;
        if (s.vt)
        {
            s.dt_idx[k] = idx;
            idx += 1;
        }
    }
    int xf_free = 0;
    for (int i = 0; i < nx; ++i) xf_free += s.xfFixed(i) ? 0 : 1;
    if (xf_free > 0)
    {
        s.x_idx[K] = idx;
        idx += xf_free;
    }
    const int n = idx;

    // ---- device layout <-> reference order -----------------------------------------------------------------------------------
    s.ref_of_internal.assign((size_t)K * s.nb, -1);
    s.internal_of_ref.assign(n, -1);
    for (int k = 0; k < K; ++k)
    {
        const int base = k * s.nb;
        for (int i = 0; i < nu; ++i) s.ref_of_internal[base + i] = s.u_idx[k] + i;
        if (s.vt) s.ref_of_internal[base + nu] = s.dt_idx[k];
        if (k + 1 < K)
            for (int i = 0; i < nx; ++i) s.ref_of_internal[base + nu + s.vt + i] = s.x_idx[k + 1] + i;
        else
        {
            int f = 0;
            for (int i = 0; i < nx; ++i)
                if (!s.xfFixed(i)) s.ref_of_internal[base + nu + s.vt + i] = s.x_idx[K] + f++;
        }
    }
    for (size_t j = 0; j < s.ref_of_internal.size(); ++j)
        if (s.ref_of_internal[j] >= 0) s.internal_of_ref[s.ref_of_internal[j]] = (int32_t)j;

    // ---- edge indices ------------------------------------------------------------------------------------------------------
    s.state_cost_idx.assign(K, -1);
    s.control_cost_idx.assign(K, -1);
    s.dt_cost_idx.assign(2 * (size_t)K, -1);
    s.dynamics_idx.assign(K, -1);
    s.dt_eq_idx.assign(K, -1);
    int lsq = 0, eq = 0;
    for (int k = 0; k < K; ++k)
    {
        if (ocp.stage_cost == B200SQP_COST_QUADRATIC_LSQ)
        {
            s.state_cost_idx[k] = lsq;
            lsq += nx;
            s.control_cost_idx[k] = lsq;
            lsq += nu;
        }
        else if (ocp.stage_cost == B200SQP_COST_MINIMUM_TIME_LSQ && (k == 0 || !single_dt))
        {
            s.dt_cost_idx[2 * k] = lsq++;      // nlp_functions.cpp:91-98
            s.dt_cost_idx[2 * k + 1] = lsq++;  // nlp_functions.cpp:100-107 (the duplicate)
        }
        s.dynamics_idx[k] = eq;
        eq += nx;
        if (s.dteq && k > 0) s.dt_eq_idx[k] = eq++;  // TwoScalarEqualEdge(dt_{k-1}, dt_k), non_uniform_finite_differences_variable_grid.cpp:150-154
    }
    if (xf_free > 0 && ocp.final_cost == 1)
    {
        s.final_cost_idx = lsq;
        lsq += nx;
    }
    // final-stage constraint edge on xf, created after all interval edges and only if xf is not fully fixed
    // (finite_differences_grid.cpp:131-144; same in the non-uniform and shooting grids)
    int ineq = 0;
    if (xf_free > 0 && ocp.final_constraint == B200SQP_FINAL_CONSTRAINT_EQUALITY)
    {
        s.final_eq_idx = eq;
        eq += nx;  // TerminalEqualityConstraint::getNonIntegralStateTermDimension = xref.size()
    }
    else if (xf_free > 0 && ocp.final_constraint == B200SQP_FINAL_CONSTRAINT_BALL)
    {
        s.final_ineq_idx = ineq;
        ineq += 1;
    }

    // ---- bounds rows: active vertices in order, unfixed components with a finite bound -----------------------------------------
    s.bound_row.assign(n, -1);
    int mb = 0;
    auto boundsOf = [&](int ref0, int dim, const double* lb, const double* ub, const int32_t* fixed) {
        int f = 0;
        for (int i = 0; i < dim; ++i)
        {
            if (fixed && fixed[i]) continue;
            if (finiteLb(lb[i]) || finiteUb(ub[i])) s.bound_row[ref0 + f] = mb++;
            ++f;
        }
    };
    for (int k = 0; k < K; ++k)
    {
        if (k > 0) boundsOf(s.x_idx[k], nx, ocp.x_lb, ocp.x_ub, nullptr);
        boundsOf(s.u_idx[k], nu, ocp.u_lb, ocp.u_ub, nullptr);
        if (s.vt) boundsOf(s.dt_idx[k], 1, &ocp.dt_lb, &ocp.dt_ub, nullptr);
    }
    if (xf_free > 0) boundsOf(s.x_idx[K], nx, ocp.x_lb, ocp.x_ub, ocp.xf_fixed);

    // ---- combined Jacobian pattern -------------------------------------------------------------------------------------------
    const int eq_start = lsq, ineq_start = lsq + eq, b_start = ineq_start + ineq;
    std::vector<std::pair<int, int>> entries;  // (col, row)
    auto block = [&](int row0, int rows, int col0, int cols) {
        if (col0 < 0) return;
        for (int c = 0; c < cols; ++c)
            for (int r = 0; r < rows; ++r) entries.push_back({col0 + c, row0 + r});
    };
    for (int k = 0; k < K; ++k)
    {
        if (s.state_cost_idx[k] >= 0) block(s.state_cost_idx[k], nx, s.x_idx[k], nx);
        if (s.control_cost_idx[k] >= 0) block(s.control_cost_idx[k], nu, s.u_idx[k], nu);
        for (int r = 0; r < 2; ++r)
            if (s.dt_cost_idx[2 * k + r] >= 0) block(s.dt_cost_idx[2 * k + r], 1, s.dt_idx[k], 1);
        const int row0 = eq_start + s.dynamics_idx[k];
        block(row0, nx, s.x_idx[k], nx);
        block(row0, nx, s.u_idx[k], nu);
        block(row0, nx, s.x_idx[k + 1], k + 1 < K ? nx : xf_free);
        block(row0, nx, s.dt_idx[k], 1);
        if (s.dt_eq_idx[k] >= 0)
        {
            block(eq_start + s.dt_eq_idx[k], 1, s.dt_idx[k - 1], 1);
            block(eq_start + s.dt_eq_idx[k], 1, s.dt_idx[k], 1);
        }
    }
    if (s.final_cost_idx >= 0) block(s.final_cost_idx, nx, s.x_idx[K], xf_free);
    if (s.final_eq_idx >= 0) block(eq_start + s.final_eq_idx, nx, s.x_idx[K], xf_free);
    if (s.final_ineq_idx >= 0) block(ineq_start + s.final_ineq_idx, 1, s.x_idx[K], xf_free);
    for (int c = 0; c < n; ++c)
        if (s.bound_row[c] >= 0) entries.push_back({c, b_start + s.bound_row[c]});
    std::sort(entries.begin(), entries.end());
    s.col_ptr.assign(n + 1, 0);
    s.row_idx.resize(entries.size());
    for (size_t e = 0; e < entries.size(); ++e)
    {
        s.col_ptr[entries[e].first + 1]++;
        s.row_idx[e] = entries[e].second;
    }
    for (int c = 0; c < n; ++c) s.col_ptr[c + 1] += s.col_ptr[c];

    // structural nnz of triu(J^T J): two columns couple iff they share a row
    const int m = lsq + eq + ineq + mb;
    int nnzH = 0;
    {
        std::vector<std::vector<int>> rows(m);
        for (auto& e : entries) rows[e.second].push_back(e.first);
        int bw = 0;
        for (auto& r : rows)
            if (!r.empty()) bw = std::max(bw, *std::max_element(r.begin(), r.end()) - *std::min_element(r.begin(), r.end()));
        std::vector<char> band((size_t)n * (bw + 1), 0);
        for (auto& r : rows)
            for (int a : r)
                for (int b : r)
                    if (b <= a) band[(size_t)a * (bw + 1) + (a - b)] = 1;
        for (char c : band) nnzH += c;
    }

    b200sqp_dims& d     = s.dims;
    d.n_params          = n;
    d.m_lsq             = lsq;
    d.m_eq              = eq;
    d.m_ineq            = ineq;
    d.m_bounds          = mb;
    d.nnz_jacobian      = (int32_t)entries.size();
    d.nnz_hessian_upper = nnzH;
    d.n_blocks          = K;
    d.block_dim         = s.nb;
    // s*[2*(nnzJ + nnzH + nnzL + 2m + 2n) + 4n], nnzL = nnzH (block tridiagonal, no fill), s = 8 (fp64)
    d.algorithmic_bytes_per_iteration = 8 * (2 * ((int64_t)d.nnz_jacobian + 2 * (int64_t)nnzH + 2 * (int64_t)m + 2 * (int64_t)n) + 4 * (int64_t)n);

    // ---- per-interval scatter tables for the materialising evaluation ---------------------------------------------------------
    EvalLayout L{nx, nu, s.vt};
    s.values_per_interval = L.v_count();
    s.jac_per_interval    = L.j_count();
    s.value_rows.assign((size_t)K * L.v_count(), -1);
    s.jac_pos.assign((size_t)K * L.j_count(), -1);
    std::map<std::pair<int, int>, int> pos;  // (col,row) -> CSC position
    for (size_t e = 0; e < entries.size(); ++e) pos[entries[e]] = (int)e;
    auto at = [&](int col, int row) -> int {
        if (col < 0 || row < 0) return -1;
        auto it = pos.find({col, row});
        return it == pos.end() ? -1 : it->second;
    };
    for (int k = 0; k < K; ++k)
    {
        int32_t* vr = &s.value_rows[(size_t)k * L.v_count()];
        int32_t* jp = &s.jac_pos[(size_t)k * L.j_count()];
        // rows
        if (k == 0 && s.state_cost_idx[0] >= 0)
            for (int i = 0; i < nx; ++i) vr[L.v_x0c() + i] = s.state_cost_idx[0] + i;
        if (s.control_cost_idx[k] >= 0)
            for (int i = 0; i < nu; ++i) vr[L.v_uc() + i] = s.control_cost_idx[k] + i;
        for (int r = 0; r < 2; ++r) vr[L.v_tc() + r] = s.dt_cost_idx[2 * k + r];
        int xs_row = -1;  // lsq edge on x_{k+1}
        if (k + 1 < K)
            xs_row = s.state_cost_idx[k + 1];
        else
            xs_row = s.final_cost_idx;
        if (xs_row >= 0)
            for (int i = 0; i < nx; ++i) vr[L.v_xs() + i] = xs_row + i;
        for (int i = 0; i < nx; ++i) vr[L.v_e() + i] = eq_start + s.dynamics_idx[k] + i;
        // columns (reference parameter index of each component, -1 if fixed)
        std::vector<int> ucol(nu), xcol(nx, -1), ncol(nx, -1);
        for (int i = 0; i < nu; ++i) ucol[i] = s.u_idx[k] + i;
        const int tcol = s.dt_idx[k];
        if (k > 0)
            for (int i = 0; i < nx; ++i) xcol[i] = s.x_idx[k] + i;
        for (int i = 0; i < nx; ++i) ncol[i] = s.ref_of_internal[(size_t)k * s.nb + nu + s.vt + i];
        for (int i = 0; i < nu; ++i)
        {
            int br             = s.bound_row[ucol[i]];
            vr[L.v_ub() + i]   = br >= 0 ? b_start + br : -1;
            jp[L.j_ub() + i]   = br >= 0 ? at(ucol[i], b_start + br) : -1;
            jp[L.j_uc() + i]   = s.control_cost_idx[k] >= 0 ? at(ucol[i], s.control_cost_idx[k] + i) : -1;
        }
        if (tcol >= 0)
        {
            int br         = s.bound_row[tcol];
            vr[L.v_tb()]   = br >= 0 ? b_start + br : -1;
            jp[L.j_tb()]   = br >= 0 ? at(tcol, b_start + br) : -1;
            for (int r = 0; r < 2; ++r) jp[L.j_tc() + r] = at(tcol, s.dt_cost_idx[2 * k + r]);
        }
        for (int i = 0; i < nx; ++i)
        {
            if (ncol[i] >= 0)
            {
                int br           = s.bound_row[ncol[i]];
                vr[L.v_xb() + i] = br >= 0 ? b_start + br : -1;
                jp[L.j_xb() + i] = br >= 0 ? at(ncol[i], b_start + br) : -1;
                if (xs_row >= 0) jp[L.j_xs() + i] = at(ncol[i], xs_row + i);
                if (xs_row >= 0)
                    for (int r = 0; r < nx; ++r) jp[L.j_xsd() + i * nx + r] = at(ncol[i], xs_row + r);
            }
        }
        if (s.control_cost_idx[k] >= 0)
            for (int c = 0; c < nu; ++c)
                for (int r = 0; r < nu; ++r) jp[L.j_ucd() + c * nu + r] = at(ucol[c], s.control_cost_idx[k] + r);
        const int erow = eq_start + s.dynamics_idx[k];
        for (int c = 0; c < nx; ++c)
            for (int r = 0; r < nx; ++r)
            {
                jp[L.j_A() + c * nx + r] = at(xcol[c], erow + r);
                jp[L.j_C() + c * nx + r] = at(ncol[c], erow + r);
            }
        for (int c = 0; c < nu; ++c)
            for (int r = 0; r < nx; ++r) jp[L.j_Bu() + c * nx + r] = at(ucol[c], erow + r);
        for (int r = 0; r < nx; ++r) jp[L.j_Bt() + r] = at(tcol, erow + r);
        if (s.dt_eq_idx[k] >= 0)
        {
            vr[L.v_dq()]     = eq_start + s.dt_eq_idx[k];
            jp[L.j_dq()]     = at(s.dt_idx[k - 1], eq_start + s.dt_eq_idx[k]);
            jp[L.j_dq() + 1] = at(tcol, eq_start + s.dt_eq_idx[k]);
        }
        if (k == K - 1)
        {
            if (s.final_eq_idx >= 0)
            {
                for (int r = 0; r < nx; ++r) vr[L.v_teq() + r] = eq_start + s.final_eq_idx + r;
                for (int c = 0; c < nx; ++c)
                    for (int r = 0; r < nx; ++r) jp[L.j_teq() + c * nx + r] = at(ncol[c], eq_start + s.final_eq_idx + r);
            }
            if (s.final_ineq_idx >= 0)
            {
                vr[L.v_tin()] = ineq_start + s.final_ineq_idx;
                for (int c = 0; c < nx; ++c) jp[L.j_tin() + c] = at(ncol[c], ineq_start + s.final_ineq_idx);
            }
        }
    }
    return B200SQP_OK;
}

}  // namespace b200sqp
