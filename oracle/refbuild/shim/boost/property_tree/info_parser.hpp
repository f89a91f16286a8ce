// TEST INFRASTRUCTURE ONLY (oracle/): a reader for the INFO files the reference ships (key value pairs, nested { } groups,
// ';' comments, optional double quotes), behind boost::property_tree::read_info's name.
#pragma once
#include "ptree.hpp"

namespace boost { namespace property_tree {

namespace info_detail {
inline std::vector<std::string> tokens(const std::string& line) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size()) {
    const char c = line[i];
    if (c == ' ' || c == '\t' || c == '\r') { ++i; continue; }
    if (c == ';') break;
    if (c == '{' || c == '}') { out.push_back(std::string(1, c)); ++i; continue; }
    if (c == '"') {
      std::string s; ++i;
      while (i < line.size() && line[i] != '"') { if (line[i] == '\\' && i + 1 < line.size()) ++i; s += line[i++]; }
      ++i; out.push_back(s); continue;
    }
    std::string s;
    while (i < line.size() && line[i] != ' ' && line[i] != '\t' && line[i] != '\r' && line[i] != ';' && line[i] != '{' && line[i] != '}') s += line[i++];
    out.push_back(s);
  }
  return out;
}
}  // namespace info_detail

inline void read_info(const std::string& filename, ptree& pt) {
  std::ifstream f(filename);
  if (!f.is_open()) throw ptree_error("cannot open file " + filename);
  pt = ptree();
  std::vector<ptree*> stack{&pt};
  ptree* last = nullptr;
  std::string line;
  while (std::getline(f, line)) {
    const std::vector<std::string> tk = info_detail::tokens(line);
    size_t i = 0;
    while (i < tk.size()) {
      if (tk[i] == "{") { if (!last) throw ptree_error("unexpected { in " + filename); stack.push_back(last); last = nullptr; ++i; continue; }
      if (tk[i] == "}") { if (stack.size() < 2) throw ptree_error("unexpected } in " + filename); stack.pop_back(); last = nullptr; ++i; continue; }
      stack.back()->kids.emplace_back(tk[i], ptree());
      last = &stack.back()->kids.back().second;
      ++i;
      if (i < tk.size() && tk[i] != "{" && tk[i] != "}") { last->value = tk[i]; ++i; }
    }
  }
  if (stack.size() != 1) throw ptree_error("unbalanced { in " + filename);
}

}}  // namespace boost::property_tree
