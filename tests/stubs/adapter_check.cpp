// compiles every member of the C++ adapter against include/lc_abi.h (tests/test_abi_cpu.py)
#include "loop_closing_b200.hpp"
int main() {
  try {
    loop_closing::b200::HammingMatcher m(0.15, 30);
    cv::Mat d;
    m.addFrame(0, d);
    auto g = m.matchFeatures(d, d);
    auto l = m.detectLoops(1, d);
    m.beginDetectLoops(1, d);
    auto l2 = m.finishDetectLoops();
    auto c = m.matchCandidates(d, {0});
    m.beginFrame(2, d);
    auto fm = m.frameMatches();
    auto fl = m.finishFrame();
    loop_closing::b200::SiftMatcher s;
    std::vector<cv::DMatch> good;
    s.matchFeatures(d, d, good, 0.75);
    s.addView(0, d);
    auto r = s.loopMatchCounts(d, 5, 3);
    auto ls = s.loopSearch(3);
    return static_cast<int>(fm.size() + fl.size() + g.size() + l.size() + l2.size() + c.size() + good.size() + r.size() + ls.size());
  } catch (const std::exception&) {
    return 0;  // no B200 here: lc_create throws, which is the point of "no CPU fallback"
  }
}
