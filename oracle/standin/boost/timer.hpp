// TEST INFRASTRUCTURE ONLY -- stand-in for the legacy <boost/timer.hpp> (boost::timer).
#ifndef RSM_STANDIN_BOOST_TIMER_HPP
#define RSM_STANDIN_BOOST_TIMER_HPP
#include <chrono>
namespace boost {
class timer {
 public:
  timer() : t0_(std::chrono::steady_clock::now()) {}
  void restart() { t0_ = std::chrono::steady_clock::now(); }
  double elapsed() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
  }
 private:
  std::chrono::steady_clock::time_point t0_;
};
}  // namespace boost
#endif
