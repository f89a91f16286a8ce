// TEST INFRASTRUCTURE ONLY (oracle/): a reader for the JSON weight files the reference ships (objects, arrays, numbers, strings, true /
// false / null), behind boost::property_tree::read_json's name. As in Boost, array elements become children with an empty key and every
// scalar is kept as its text.
#pragma once
#include <cctype>
#include <iterator>
#include "ptree.hpp"

namespace boost { namespace property_tree {

namespace json_detail {
struct Parser {
  const std::string& s; size_t i = 0; const std::string& file;
  Parser(const std::string& text, const std::string& f) : s(text), file(f) {}
  [[noreturn]] void fail(const char* what) { throw ptree_error(std::string("json: ") + what + " in " + file); }
  void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  std::string str() {
    std::string out; ++i;
    while (i < s.size() && s[i] != '"') { if (s[i] == '\\' && i + 1 < s.size()) ++i; out += s[i++]; }
    if (i >= s.size()) fail("unterminated string");
    ++i; return out;
  }
  void value(ptree& t) {
    ws();
    if (i >= s.size()) fail("unexpected end");
    if (s[i] == '{') {
      ++i; ws();
      if (s[i] == '}') { ++i; return; }
      for (;;) {
        ws(); if (s[i] != '"') fail("key expected");
        const std::string key = str();
        ws(); if (s[i] != ':') fail(": expected"); ++i;
        t.kids.emplace_back(key, ptree());
        value(t.kids.back().second);
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == '}') { ++i; return; }
        fail(", or } expected");
      }
    } else if (s[i] == '[') {
      ++i; ws();
      if (s[i] == ']') { ++i; return; }
      for (;;) {
        t.kids.emplace_back(std::string(), ptree());
        value(t.kids.back().second);
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == ']') { ++i; return; }
        fail(", or ] expected");
      }
    } else if (s[i] == '"') {
      t.value = str();
    } else {
      size_t j = i;
      while (j < s.size() && s[j] != ',' && s[j] != '}' && s[j] != ']' && !std::isspace((unsigned char)s[j])) ++j;
      if (j == i) fail("value expected");
      t.value = s.substr(i, j - i);
      i = j;
    }
  }
};
}  // namespace json_detail

inline void read_json(const std::string& filename, ptree& pt) {
  std::ifstream f(filename);
  if (!f.is_open()) throw ptree_error("cannot open file " + filename);
  const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  pt = ptree();
  json_detail::Parser p(text, filename);
  p.value(pt);
}

}}  // namespace boost::property_tree
