// smokeSimulation.cuh -- drop-in replacement for the reference header of the same name
// (reference: project/smokeSimulation.cuh:1-18).  Same nine C++-linkage entry points, same argument
// meaning, same blocking behaviour, so the reference's main.cpp / boundingBox.cpp compile and link
// against libsmoke_b200.so unchanged.  Implemented in csrc/dropin.cu as thin
// forwards to the C ABI in include/smoke_b200.h on one process-global handle.
#pragma once
#include <vector>

// prints one block of properties per visible CUDA device (reference cu:62-85)
void getGPUProperties(void);

// one simulation step; smoke_grid (W*H*D floats, x fastest) receives the new density before returning
// (reference cu:774-819).  Passing nullptr skips the device->host copy (extension).
void simulate(float* smoke_grid, float dt);

// allocate the volume; smoke_grid holds the initial density (reference cu:124-238)
void initializeVolume(float* smoke_grid, unsigned int width, unsigned int heigth, unsigned int depth);
void deleteVolume();

// scene objects; ids are dense and shared between obstacles and sources (reference cu:88-109)
int addObstacle(float x, float y, float z, float vx, float vy, float vz, float r);
int addSmokeSource(float x, float y, float z, float r);
void updateObjectPos(int id, float x, float y, float z);

// stable pointers the GUI writes through between steps (reference cu:31-36)
float* getBuoyancy();
float* getGravity();
