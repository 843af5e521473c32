// Thermostatted / sheared integrators and the multi-GPU step phases (filled in after the NVE path).
#include "common.cuh"
#include "integrate.cuh"
#include "sim.cuh"

using namespace atlj;

#define NOT_YET(name)                                   \
  do {                                                  \
    set_error(name ": not implemented in this build");  \
    return ATLJ_ERR_ARG;                                \
  } while (0)

extern "C" {
int atlj_sim_run_baoab(atlj_sim *, double, double, double, uint64_t, int, double *) { NOT_YET("atlj_sim_run_baoab"); }
int atlj_sim_run_baoab_noise(atlj_sim *, double, double, double, const double *, int, double *) {
  NOT_YET("atlj_sim_run_baoab_noise");
}
int atlj_sim_run_nhc(atlj_sim *, double, double, int, double *, double *, const double *, int, double *) {
  NOT_YET("atlj_sim_run_nhc");
}
int atlj_sim_run_sllod(atlj_sim *, double, double, double *, int, double *) { NOT_YET("atlj_sim_run_sllod"); }
int atlj_sim_mg_drift_pack(atlj_sim *, double, double *, double *, int64_t) { NOT_YET("atlj_sim_mg_drift_pack"); }
int atlj_sim_mg_bin_pack(atlj_sim *, const double *, const double *, int64_t, double *, double *, int64_t) {
  NOT_YET("atlj_sim_mg_bin_pack");
}
int atlj_sim_mg_force_kick(atlj_sim *, const double *, const double *, int64_t, double, double *) {
  NOT_YET("atlj_sim_mg_force_kick");
}
}
