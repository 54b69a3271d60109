/* stub: see Rinternals.h in this directory */
#include "Rinternals.h"
