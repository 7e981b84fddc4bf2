#pragma once
#include "core.h"
