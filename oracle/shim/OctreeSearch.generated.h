// TEST INFRASTRUCTURE ONLY. UnrealHeaderTool would generate this (OctreeSearch.h:6); nothing is needed.
#pragma once
