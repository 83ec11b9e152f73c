#pragma once
// Stand-in for RcppProgress.  The tests can simulate a user interrupt: check_abort() answers true
// once abort_after() calls have been made (negative: never).
class Progress {
public:
  Progress(int, bool) {}
  static int &abort_after() { static int n = -1; return n; }
  static int &increments() { static int n = 0; return n; }
  static bool check_abort() {
    int &n = abort_after();
    if (n < 0) return false;
    if (n == 0) return true;
    n--;
    return false;
  }
  void increment() { increments()++; }
};
