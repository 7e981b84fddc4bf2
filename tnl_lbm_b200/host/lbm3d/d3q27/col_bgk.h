// forwards to the host mirror of the reference interface (trait classes are tags here, see ../lbmx_host.h)
#pragma once
#include "../lbmx_host.h"
