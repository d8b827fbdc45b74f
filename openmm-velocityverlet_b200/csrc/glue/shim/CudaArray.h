#include "openmm_shim_all.h"   // compile-check stand-in, see openmm_shim_all.h
