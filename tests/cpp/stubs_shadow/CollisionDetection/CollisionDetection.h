// Lock-step shadow: the engine's sources see ONE class CollisionDetection that feeds every entry to both the reference's own class
// (renamed ReferenceCollisionDetection; its CollisionDetection.cpp is compiled with the same -D rename) and the drop-in over libimrcd.so.
// The reference's callbacks go to the ECS and drive the game; the drop-in's callbacks for the SAME entries go to a sink and are compared
// frame by frame (tests/cpp/snake_harness.cpp).  Same inputs every frame, so no trajectory divergence can hide or fake a difference.
#pragma once
#include <unordered_map>
#include <utility>
#include <vector>

#include "ECS/ECStypes.h"
#include "ECS/ECSwrapper.h"
#include "Geometry/OBBtree.h"
#ifndef IMRCD_RECORD_ONLY
#include "imrcd_host.hpp"
#endif

#define CollisionDetection ReferenceCollisionDetection
#include_next "CollisionDetection/CollisionDetection.h"
#undef CollisionDetection

void imrcd_shadow_sink(const std::vector<std::pair<Entity, std::vector<CollisionCallbackData>>>& callbacks);     // snake_harness.cpp
void imrcd_shadow_compare();
void imrcd_shadow_entry(const CollisionDetectionEntry& entry);

#ifndef IMRCD_RECORD_ONLY        // -DIMRCD_RECORD_ONLY: the reference alone behind the tee (CPU only), for writing frame dumps
namespace imrcd_shadow {
#define IMRCD_DROP_IN_CALLBACK_SINK(callbacks) ::imrcd_shadow_sink(callbacks)
#include "CollisionDetection_drop_in.hpp"
}
#endif

class CollisionDetection
{
public:
#ifdef IMRCD_RECORD_ONLY
    CollisionDetection(ECSwrapper* in_ECSwrapper_ptr) : reference(in_ECSwrapper_ptr) {}
    void Reset() { reference.Reset(); }
#else
    CollisionDetection(ECSwrapper* in_ECSwrapper_ptr) : reference(in_ECSwrapper_ptr), drop_in(in_ECSwrapper_ptr) {}
    void Reset() { reference.Reset(); drop_in.Reset(); }
#endif
    void AddCollisionDetectionEntry(const CollisionDetectionEntry in_collisionDetectionEntry)
    {
        imrcd_shadow_entry(in_collisionDetectionEntry);
        reference.AddCollisionDetectionEntry(in_collisionDetectionEntry);
#ifndef IMRCD_RECORD_ONLY
        drop_in.AddCollisionDetectionEntry(in_collisionDetectionEntry);
#endif
    }
    void ExecuteCollisionDetection()
    {
#ifndef IMRCD_RECORD_ONLY
        drop_in.ExecuteCollisionDetection();        // -> imrcd_shadow_sink
#endif
        reference.ExecuteCollisionDetection();      // -> the ECS components (the game moves by the reference's deltaVectors)
        imrcd_shadow_compare();
    }
private:
    ReferenceCollisionDetection reference;
#ifndef IMRCD_RECORD_ONLY
    imrcd_shadow::CollisionDetection drop_in;
#endif
};
