/* TEST INFRASTRUCTURE ONLY: scope for the per-node loop of the reference's phase 2 (cpp/exec/psp_process.cpp:2460-2498: gain,
 * Iref/I, detrend, delta pressure, delta Cp, rms / avg partial sums) when it is compiled on its own as the body of
 * ref_phase2_nodes (see the _ref/phase2.o rule of the Makefile).  The reference's own paint calibration and tunnel conditions
 * (cpp/include/non_cv_upsp.h) are used as they are; its TransPolyFitter needs Eigen's QR, which is absent, so `pfitter` here
 * hands the series to the oracle's restatement of that solve (orc_transpoly_eval_fit, resolved by the probe with dlopen):
 * everything around the fit is the reference's compiled arithmetic. */
#include <cmath>
#include <limits>
#include <vector>
#include "non_cv_upsp.h"
struct OracleFitter {
  void (*eval)(const float* A, unsigned n_frames, unsigned ncoef, const float* data, float* fit, float* coef_out, float* scratch);
  std::vector<float> A;
  unsigned n_frames, ncoef;
  void skip_fit(unsigned int) {}
  std::vector<float> eval_fit(float* data, unsigned int, unsigned int) {
    std::vector<float> fit(n_frames), scratch((size_t)n_frames * (ncoef + 1));
    eval(A.data(), n_frames, ncoef, data, fit.data(), nullptr, scratch.data());
    return fit;
  }
};
void ref_phase2_nodes(unsigned int my_num_nodes, unsigned int node_start, unsigned int number_frames, const std::vector<float>& coverage,
                      OracleFitter& pfitter, std::vector<double>& local_rms, std::vector<double>& local_avg, std::vector<double>& local_gain,
                      upsp::TunnelConditions& tcond, const std::vector<float>& steady, upsp::PaintCalibration& pcal,
                      const std::vector<float>& model_temp_input, const std::vector<float>& sol_avg_final,
                      const float* ptr_intensity_transpose_data, float* ptr_pressure_transpose_data);
