/* stand-in for R.h (TEST HARNESS ONLY, see Rinternals.h) */
#include "Rinternals.h"
