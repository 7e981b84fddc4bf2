// defs.h -- forwards to the host mirror of the reference interface (see lbmx_host.h)
#pragma once
#include "lbmx_host.h"
