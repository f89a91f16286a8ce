// info_reader.h — tiny readers for the reference's settings files, which stay byte-identical:
//   * Boost.PropertyTree INFO  (ddp_setting.info, constraint_params*.info, mhpc_config.info;
//     read in the reference by loadHSDDPSetting, HSDDP_CompoundTypes.h:57-82, load_reb_params /
//     load_al_params, ConstraintsBase.h:88-111, loadConstrintParameters, HKDProblem.h:70-90,
//     loadMHPCConfig, MHPCProblem.h:67-83)
//   * the JSON cost-weight files (loadCostWeights, MHPCCostUtil.h:10-143)
// Boost is not available in this image; keys and semantics are the same ("section.key").
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cafe {

class InfoFile {
 public:
  explicit InfoFile(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open settings file " + path);
    std::vector<std::string> tok;
    std::string line;
    while (std::getline(f, line)) {
      size_t sc = line.find(';');  // ';' starts a comment in INFO files
      if (sc != std::string::npos) line = line.substr(0, sc);
      std::istringstream ls(line);
      std::string w;
      std::vector<std::string> words;
      while (ls >> w) words.push_back(w);
      for (auto& x : words) tok.push_back(x);
      tok.push_back("\n");
    }
    std::vector<std::string> prefix;
    std::string pending;  // key waiting for '{' or a value
    bool eol = false;  // a key alone on its line may still open a section on the next line
    for (size_t i = 0; i < tok.size(); ++i) {
      const std::string& t = tok[i];
      if (t == "\n") { eol = true; continue; }
      if (t == "{") { prefix.push_back(pending); pending.clear(); eol = false; continue; }
      if (t == "}") { if (!prefix.empty()) prefix.pop_back(); pending.clear(); eol = false; continue; }
      if (pending.empty() || eol) { pending = t; eol = false; continue; }
      std::string key;
      for (auto& p : prefix) key += p + ".";
      kv_[key + pending] = t;
      pending.clear();
    }
  }
  bool has(const std::string& key) const { return kv_.count(key) > 0; }
  std::string str(const std::string& key) const {
    auto it = kv_.find(key);
    if (it == kv_.end()) throw std::runtime_error("missing setting " + key);
    return it->second;
  }
  double num(const std::string& key) const { return std::strtod(str(key).c_str(), nullptr); }
  int integer(const std::string& key) const { return (int)std::strtol(str(key).c_str(), nullptr, 10); }
  bool boolean(const std::string& key) const { std::string s = str(key); return s == "true" || s == "1"; }

 private:
  std::map<std::string, std::string> kv_;
};

// Flat JSON reader: {"Section": {"key": number | [numbers]}} -> "Section.key" -> vector<double>
class JsonWeights {
 public:
  explicit JsonWeights(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open cost file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    s_ = ss.str();
    size_t i = 0;
    skip(i);
    if (s_[i] != '{') throw std::runtime_error("bad JSON " + path);
    ++i;
    while (true) {
      skip(i);
      if (s_[i] == '}') break;
      std::string sec = qstr(i);
      skip(i); expect(i, ':'); skip(i); expect(i, '{');
      while (true) {
        skip(i);
        if (s_[i] == '}') { ++i; break; }
        std::string key = qstr(i);
        skip(i); expect(i, ':'); skip(i);
        std::vector<double> v;
        if (s_[i] == '[') {
          ++i;
          while (true) { skip(i); if (s_[i] == ']') { ++i; break; } v.push_back(number(i)); skip(i); if (s_[i] == ',') ++i; }
        } else v.push_back(number(i));
        kv_[sec + "." + key] = v;
        skip(i);
        if (s_[i] == ',') ++i;
      }
      skip(i);
      if (s_[i] == ',') ++i;
    }
  }
  const std::vector<double>& vec(const std::string& key) const {
    auto it = kv_.find(key);
    if (it == kv_.end()) throw std::runtime_error("missing cost weight " + key);
    return it->second;
  }
  double num(const std::string& key) const { return vec(key).at(0); }

 private:
  void skip(size_t& i) { while (i < s_.size() && std::isspace((unsigned char)s_[i])) ++i; }
  void expect(size_t& i, char c) { if (s_[i] != c) throw std::runtime_error(std::string("JSON: expected ") + c); ++i; }
  std::string qstr(size_t& i) { expect(i, '"'); size_t j = s_.find('"', i); std::string r = s_.substr(i, j - i); i = j + 1; return r; }
  double number(size_t& i) { char* e; double v = std::strtod(s_.c_str() + i, &e); i = (size_t)(e - s_.c_str()); return v; }
  std::string s_;
  std::map<std::string, std::vector<double>> kv_;
};

}  // namespace cafe
