/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's LM/SQP hot path.  See sqp_oracle.cpp. */
#ifndef SQP_ORACLE_H_
#define SQP_ORACLE_H_

#include "../include/b200sqp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* > 1: every `xref` argument below is a time-varying reference [n_grid][nx] (row k = getReferenceCached(k)); 0 / 1: static [nx] */
int sqp_oracle_set_xref_points(int n_points);
int sqp_oracle_dims(const b200sqp_ocp* d, b200sqp_dims* out);
int sqp_oracle_vertex_indices(const b200sqp_ocp* d, int32_t* x_idx, int32_t* u_idx, int32_t* dt_idx);
int sqp_oracle_edge_table(const b200sqp_ocp* d, int category, int32_t* table, int max_edges);
int sqp_oracle_initial_params(const b200sqp_ocp* d, const double* x0, const double* xref, double* params);
int sqp_oracle_evaluate(const b200sqp_ocp* d, const double* x0, const double* xref, const double* params, double w_eq, double w_ineq, double w_b,
                        double* values, double* jac_dense, uint8_t* jac_pattern, double* params_after);
int sqp_oracle_trace(const b200sqp_ocp* d, const b200sqp_lm_options* o, const double* x0, const double* xref, const double* params_in,
                     double* params_out, double* chi2_out, int32_t* status_out, int max_events, int32_t* ev_type, double* ev_chi2, double* ev_vec,
                     int32_t* n_events);
int sqp_oracle_solve_batch(const b200sqp_ocp* d, const b200sqp_lm_options* o, int batch, const double* x0, const double* xref,
                           const double* params_in, double* params_out, double* chi2, int32_t* status, int threads, double* seconds);

int sqp_oracle_solve_sequence(const b200sqp_ocp* d, const b200sqp_lm_options* o, const double* x0, const double* xref, int n_solves,
                              double* params_out, double* chi2_out);
int sqp_oracle_plant_step(const b200sqp_ocp* d, int integrator, double dt, int batch, const double* x, const double* u, double* x_next);
double sqp_oracle_plant_interval(double plant_dt, int step);
int sqp_oracle_dynamics_hessian(const b200sqp_ocp* d, int method, const double* x0, const double* u0, const double* multipliers, double* H);
int sqp_oracle_linearize(const b200sqp_ocp* d, int method, const double* x0, const double* u0, double* A, double* B);
int sqp_oracle_known_answer(int case_id, int stage, double* x_out, double* expected, double* tol, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
