// TEST INFRASTRUCTURE (oracle/_ref build only). Not part of the product library.
//
// The reference defines these member functions only in viewport.cpp (the GLUT/OpenGL
// preview, reference BHRayTracer/viewport.cpp:545-685), which is out of scope and is not
// compiled into the headless oracle.  Their vtables still need bodies to link.
#include "Scenes/scene.h"
#include "Objects/objects.h"
#include "Materials/materials.h"
#include "Lights/lights.h"
#include "Textures/texture.h"

void Sphere::ViewportDisplay(const Material *) const {}
void Plane::ViewportDisplay(const Material *) const {}
void TriObj::ViewportDisplay(const Material *) const {}
void MtlBlinn::SetViewportMaterial(int) const {}
void GenLight::SetViewportParam(int, ColorA, ColorA, Vec4f) const {}
void PointLight::SetViewportLight(int) const {}
bool TextureFile::SetViewportTexture() const { return false; }
bool TextureChecker::SetViewportTexture() const { return false; }

void BeginRender();
// reference Main.cpp:428 calls ShowViewport() from its main(); that main is renamed away
// (-Dmain=ref_main_unused) and the driver calls BeginRender() itself.
void ShowViewport() { BeginRender(); }
