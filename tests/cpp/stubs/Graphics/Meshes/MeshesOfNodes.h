// Test double for IMR/include/Graphics/Meshes/MeshesOfNodes.h (the real one needs the Vulkan / VMA headers): just what
// ModelCollisionCompEntity::AddCollisionDetectionEntryToVector reads, MeshInfo::boundBoxTree (ModelCollisionCompEntity.cpp:74).
#pragma once
#include <vector>
#include "Geometry/OBBtree.h"
struct MeshInfo { OBBtree boundBoxTree; };
class MeshesOfNodes {
public:
    const MeshInfo& GetMeshInfo(size_t index) const { return meshes[index]; }
    std::vector<MeshInfo> meshes;
};
