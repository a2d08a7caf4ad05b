/* Part of the GSL stand-in used only to build oracle/_ref (see gsl_linalg.h). */
#include "gsl_linalg.h"
