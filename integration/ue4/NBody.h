// Module header of the adapted game module: same role as the reference's Source/NBody/NBody.h (which pulls in the engine
// umbrella header for every file of the module).
#pragma once
#include "Engine.h"
