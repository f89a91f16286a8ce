// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the part of boost::property_tree the reference's settings loaders use
// (read_info + ptree::get<T>("a.b"), HSDDP_CompoundTypes.h:57-82, HKDProblem.h:68-90). Boost is not in this image.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace boost { namespace property_tree {

class ptree_error : public std::runtime_error { public: explicit ptree_error(const std::string& w) : std::runtime_error(w) {} };
class ptree_bad_path : public ptree_error { public: explicit ptree_bad_path(const std::string& w) : ptree_error(w) {} };
class ptree_bad_data : public ptree_error { public: explicit ptree_bad_data(const std::string& w) : ptree_error(w) {} };

class ptree {
 public:
  std::string value;
  std::vector<std::pair<std::string, ptree>> kids;
  typedef std::vector<std::pair<std::string, ptree>>::iterator iterator;
  typedef std::vector<std::pair<std::string, ptree>>::const_iterator const_iterator;
  iterator begin() { return kids.begin(); }
  iterator end() { return kids.end(); }
  const_iterator begin() const { return kids.begin(); }
  const_iterator end() const { return kids.end(); }
  std::size_t size() const { return kids.size(); }
  bool empty() const { return kids.empty(); }

  const ptree* find_path(const std::string& path) const {
    const ptree* n = this;
    size_t pos = 0;
    while (pos <= path.size()) {
      const size_t dot = path.find('.', pos);
      const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      const ptree* next = nullptr;
      for (const auto& kv : n->kids) if (kv.first == key) { next = &kv.second; break; }
      if (!next) return nullptr;
      n = next;
      if (dot == std::string::npos) break;
      pos = dot + 1;
    }
    return n;
  }
  const ptree& get_child(const std::string& path) const {
    const ptree* n = find_path(path);
    if (!n) throw ptree_bad_path("No such node (" + path + ")");
    return *n;
  }
  ptree& get_child(const std::string& path) { return const_cast<ptree&>(static_cast<const ptree*>(this)->get_child(path)); }
  template <class T> T get_value() const { return convert<T>(value); }
  template <class T> T get(const std::string& path) const { return get_child(path).template get_value<T>(); }
  template <class T> T get(const std::string& path, const T& dflt) const { const ptree* n = find_path(path); return n ? n->template get_value<T>() : dflt; }

 private:
  template <class T> static T convert(const std::string& s) {
    if constexpr (std::is_same<T, std::string>::value) return s;
    else if constexpr (std::is_same<T, bool>::value) {
      if (s == "true" || s == "1") return true;
      if (s == "false" || s == "0") return false;
      throw ptree_bad_data("conversion of data to type bool failed: " + s);
    } else {
      std::istringstream is(s);
      T v{};
      is >> v;
      if (is.fail()) throw ptree_bad_data("conversion of data failed: " + s);
      return v;
    }
  }
};

}}  // namespace boost::property_tree
