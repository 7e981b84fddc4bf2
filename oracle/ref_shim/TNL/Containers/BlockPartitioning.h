#pragma once
#include "NDArray.h"
