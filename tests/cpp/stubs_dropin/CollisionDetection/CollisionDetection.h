// Put in front of the engine's include path, this is the whole integration: every engine source that includes
// "CollisionDetection/CollisionDetection.h" (ModelCollisionComp.h:8, Engine.h) gets the drop-in class instead.
#pragma once
#include "CollisionDetection_drop_in.hpp"
