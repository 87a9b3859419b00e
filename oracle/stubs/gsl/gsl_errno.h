#include "gsl_rng.h"
