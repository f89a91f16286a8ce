// QuadReference.h — the handle on the top-level reference file that set-up code passes around (Reference/QuadReference.h:159-207).
// The reference object parses the CSV itself and the problem classes query it through std::function costs; here the CSV is parsed by
// the deck builders behind the C ABI (csrc/host/quad_reference.cpp: same float arithmetic, same re-ordering), so this object records
// WHICH file, whether the HKD leg order applies, and how far step() has advanced (the start offset k0 of the next deck).
#pragma once
#include <cmath>
#include <string>

class QuadReference {
 public:
  QuadReference() {}
  void load_top_level_data(const std::string& fname, bool reorder = false) { fname_ = fname; reorder_ = reorder; k_cur = 0; t_cur = 0; }
  void initialize(float plan_horizon) { dur = plan_horizon; }
  // QuadReference.cpp:33-52: k_cur advances by the number of reference rows (dt = 0.01 s in every shipped file) within dt_sim
  void step(float dt_sim) {
    for (int i = 1; (i * dt < dt_sim) || std::fabs(i * dt - dt_sim) <= 1e-6f; i++) { k_cur++; t_cur += dt; }
  }
  float get_dt() { return dt; }
  float get_start_time() { return t_cur; }
  float get_end_time() { return t_cur + dur; }
  // ---- binding
  const std::string& cafe_file() const { return fname_; }
  bool cafe_reorder() const { return reorder_; }
  int cafe_k0() const { return k_cur; }

 private:
  std::string fname_;
  bool reorder_ = false;
  int k_cur = 0;
  float t_cur = 0, dur = 0;
  float dt = 0.01f;   // rows of the reference CSVs are 10 ms apart (Reference/Data/*/quad_reference.csv, column 0)
};
