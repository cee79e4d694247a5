// Test infrastructure (oracle/): BOOST_LOG_TRIVIAL as a sink that discards everything (Boost is not installed).
#ifndef XREF_BOOST_LOG_TRIVIAL
#define XREF_BOOST_LOG_TRIVIAL
#include <ostream>
namespace xref {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}  // namespace xref
#define BOOST_LOG_TRIVIAL(lvl) ::xref::NullStream()
#endif
