#pragma once
// Stand-in for RcppProgress (only needed if sample.cpp were compiled; kept for completeness).
class Progress {
public:
  Progress(int, bool) {}
  static bool check_abort() { return false; }
  void increment() {}
};
