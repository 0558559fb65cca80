#define _GNU_SOURCE
#include <execinfo.h>
#include <dlfcn.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
static void bt(const char* what, int code) {
    void* b[64]; int n = backtrace(b, 64);
    dprintf(2, "%s(%d) called\n", what, code);
    backtrace_symbols_fd(b, n, 2);
}
void exit(int code) { bt("exit", code); void (*real)(int) = dlsym(RTLD_NEXT, "exit"); real(code); __builtin_unreachable(); }
void _exit(int code) { bt("_exit", code); void (*real)(int) = dlsym(RTLD_NEXT, "_exit"); real(code); __builtin_unreachable(); }
void _Exit(int code) { bt("_Exit", code); void (*real)(int) = dlsym(RTLD_NEXT, "_Exit"); real(code); __builtin_unreachable(); }
