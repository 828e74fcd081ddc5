/* path B oracle - filled in below */
#include "gc_oracle.h"
