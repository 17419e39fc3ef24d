// TEST INFRASTRUCTURE ONLY. UE 4.9 "GameFramework/Actor.h" stand-in (OctreeSearch.h:5); see ../Engine.h.
#pragma once
#include "Engine.h"
