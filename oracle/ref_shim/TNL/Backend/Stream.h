#pragma once
#include "Macros.h"
