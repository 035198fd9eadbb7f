// Stand-in for <boost/algorithm/string.hpp> — TEST INFRASTRUCTURE ONLY.  The reference's
// src/base/camera_models.cc includes this header without using it; src/util/string.cc calls
// boost::split(elems, str, boost::is_any_of(delim), boost::token_compress_on) in StringSplit.
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace boost {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of {
  std::string set;
  explicit is_any_of(const std::string& s) : set(s) {}
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
template <class Pred>
void split(std::vector<std::string>& out, const std::string& in, Pred pred,
           token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool last_was_delim = false;
  for (char c : in) {
    if (pred(c)) {
      if (!(mode == token_compress_on && last_was_delim)) {
        out.push_back(cur);
        cur.clear();
      }
      last_was_delim = true;
    } else {
      cur.push_back(c);
      last_was_delim = false;
    }
  }
  out.push_back(cur);
}
}  // namespace boost
