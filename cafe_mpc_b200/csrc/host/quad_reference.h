// quad_reference.h — host-side reference-trajectory source, mirroring the interface of the
// reference's QuadReference (Reference/QuadReference.h:139-207, QuadReference.cpp:6-31, :70-127,
// :134-407): same method names, same `float` time arithmetic, same float-rounded values
// (std::stof in the reference, QuadReference.cpp:159-333), same leg / body-state re-ordering.
// Only the pieces the phase-deck builders need are here; the sliding-window step() used by the
// MPC loop is SURVEY.md §8(f) "next".
#pragma once
#include <cmath>
#include <string>
#include <vector>

namespace cafe {

struct QuadAugmentedState {     // QuadReference.h:15-40
  double body_state[12] = {0};  // after load: [pos, eul, vWorld, eulrate]
  double qJ[12] = {0}, qJd[12] = {0};
  double foot_placements[12] = {0}, foot_velocities[12] = {0};
  double foot_heights[4] = {0};
  double grf[12] = {0}, torque[12] = {0};
  int contact[4] = {0};
  double status_dur[4] = {0};
};

class QuadReference {
 public:
  // reorder=true: HKD leg order FR,FL,HR,HL (QuadReference.cpp:373-407); k0 drops the first k0
  // records of the file so that a later start time runs through the untouched code path.
  void load_top_level_data(const std::string& fname, bool reorder = false, int k0 = 0);
  void initialize(float plan_horizon);                          // QuadReference.cpp:6-31
  const QuadAugmentedState* get_a_reference_ptr_at_t(float t) const;  // :70-85
  void get_contact_at_t(int contact[4], float t) const;         // :91-105
  void get_contact_duration_at_t(double dur[4], float t) const; // :110-124
  int index_at_t(float t) const;
  int get_data_size() const { return sz; }
  float get_dt() const { return dt; }
  size_t top_level_size() const { return tp_data.size(); }

 private:
  std::vector<QuadAugmentedState> tp_data, data;
  float tp_dt = 0, dt = 0, dur = 0;
  int sz = 0;
};

}  // namespace cafe
