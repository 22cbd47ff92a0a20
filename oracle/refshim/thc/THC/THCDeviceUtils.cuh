// TEST/BENCH INFRASTRUCTURE ONLY — see THC.h in this directory.
#pragma once
#include "THC.h"
