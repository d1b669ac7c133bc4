/* LD_PRELOAD helper (round-2 diagnostic): print a native backtrace on SIGSEGV. */
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void handler(int sig) {
  void* bt[96];
  int n = backtrace(bt, 96);
  (void)sig;
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
__attribute__((constructor)) static void init(void) { signal(SIGSEGV, handler); }
