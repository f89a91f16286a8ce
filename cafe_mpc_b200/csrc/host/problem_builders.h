// problem_builders.h — host-side mirrors of the reference's problem builders. They walk the
// reference contact schedule with the reference's own (float) arithmetic and produce the plain-data
// phase deck (include/cafe_deck.h) instead of a deque of SinglePhase objects wired with callbacks.
//   HKDProblem<T>   /root/reference/HKDMPC/HKD-TrajOpt/HKDProblem.{h,cpp}
//   MHPCProblem<T>  /root/reference/MHPC/MHPC-Trajopt/MHPCProblem.{h,cpp}
#pragma once
#include <string>
#include <vector>
#include "../../../include/cafe_deck.h"
#include "quad_reference.h"

namespace cafe {

// approx_eq_scalar & friends (HSDDP_Utils.h:46-78): float tolerance 1e-6, float error
template <typename T1, typename T2>
inline bool approx_eq_scalar(T1 n1, T2 n2) { float tol = 1e-6; float err = std::abs(n1 - n2); return err <= tol; }
template <typename T1, typename T2>
inline bool approx_leq_scalar(T1 n1, T2 n2) { return n1 < n2 || approx_eq_scalar(n1, n2); }
template <typename T1, typename T2>
inline bool approx_geq_scalar(T1 n1, T2 n2) { return n1 > n2 || approx_eq_scalar(n1, n2); }

struct DeckStorage {  // owns the memory a CafeDeck points into
  CafeDeck deck;
  std::vector<double> ref;
  std::vector<float> phase_start_times, phase_end_times;
};

struct HKDPlanConfig {  // HKDProblem.h:20-25
  float plan_duration;
  float timeStep;
  int nsteps_between_mpc;
};

class HKDProblem {  // HKDProblem.h:93-168
 public:
  void set_problem_data(QuadReference* quad_ref, const HKDPlanConfig& config, const std::string& constraint_params_fname);
  void initialization(DeckStorage& out);  // HKDProblem.cpp:15-111

 private:
  QuadReference* quad_ref_ptr = nullptr;
  float plan_duration = 0, dt_sim = 0, dt_mpc = 0;
  int nsteps_between_mpc = 0;
  CafeRebParam grf_reb_param{}, swing_reb_param{};
  CafeAlParam td_al_param{};
};

struct MHPCConfig {  // MHPCProblem.h:23-65: the four plan durations / time steps are doubles, dt_mpc and BG_alpha floats
  double plan_dur_wb, plan_dur_srb, dt_wb, dt_srb;
  float dt_mpc;
  double BG_alpha;     // (double)(float) of the file's value
  int num_threads;
  std::string referenceFileName, costFileName, constraintParamFileName;
};
void loadMHPCConfig(const std::string& fname, MHPCConfig& config);  // MHPCProblem.h:67-83

class MHPCProblem {  // MHPCProblem.h:169-289
 public:
  void set_problem_data(QuadReference* quad_ref, const MHPCConfig& config, const std::string& settings_root);
  void initialization(DeckStorage& out);  // MHPCProblem.cpp:13-250
  // LocoProblem (MHPC/MHPC-Trajopt/Locomotion/LocoProblem.cpp:7-84): the same builder with initialize_parameters reading only the
  // GRF / Torque / TD blocks and create_problem_one_phase attaching only the torque and GRF barriers to a whole-body phase
  bool loco = false;

 private:
  QuadReference* quad_reference = nullptr;
  MHPCConfig pconfig{};
  std::string root;
  float plan_dur_all = 0;
};

// In-place barrel roll, BarrelRollTO.cpp:65-275 (six hand-scheduled whole-body phases, fixed desired states, per-phase weights)
void build_barrel_to_deck(const std::string& cost_json, const std::string& constraint_info, DeckStorage& out);
void barrel_to_guess(const CafeDeck& deck, const double* x0, double* guess);

// Structural non-zero patterns of the HKD linearisation as 576-bit masks (bit i + 24 j; 9 words each, in the order A, B, lxx, luu):
// A and B are the CCS patterns of the generated hkinodyn_par (recorded by running it with index-collecting store functors), lxx and
// luu follow HKDModel::lq_knot (model_hkd.cuh). The backward sweep fetches only these entries.
void hkd_lq_patterns(unsigned long long out[36]);
// Structural non-zero pattern of the whole-body lxx of one running knot as a 1296-bit mask (bit i + 36 j, 21 words), from the contact
// flags of the knot's reference record `rec` (CAFE_REF_W doubles): the rule of WBModel::lq_knot (model_wb.cuh) = MHPCCost.cpp's cost
// objects: the diagonal, the base block {3,4,5,18..23}^2 shared by the feet, and per foot {3,4,5, own leg q}^2 in stance or
// {3,4,5, own leg q, 18..23, own leg v}^2 in swing.
void wb_lxx_pattern(const double* rec, unsigned long long out[21]);
// compute_hkd_state (HKDModel.h:66-96): qdummy from joint angles (swing) or foot FK (stance)
void compute_hkd_state(const double eul[3], const double pos[3], const double qJ[12], double qdummy[12], const int contact[4]);

void load_hsddp_setting(const std::string& fname, CafeOptions& o);  // loadHSDDPSetting, HSDDP_CompoundTypes.h:57-82

}  // namespace cafe
