/* TEST INFRASTRUCTURE ONLY: scope for the model-temperature lines of the reference's main() (cpp/exec/psp_process.cpp:2287-2310)
 * when they are compiled on their own (see the _ref/modeltemp.o rule of the Makefile, which also feeds the compiler the
 * reference's three constants, :1096-1098, and their member declarations, :1122-1124): the tunnel conditions they read
 * (cpp/include/non_cv_upsp.h) and a silent LOG_INFO. */
#include <cmath>
#include "non_cv_upsp.h"
#define LOG_INFO(...) ((void)0)
