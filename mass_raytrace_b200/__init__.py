"""mass-raytrace B200 backend: sm_100a wavefront path tracer behind a C ABI (include/mrt.h), with a host-side
mirror of the reference's scene API (api.py, libmrt_host.so). See DESIGN.md."""
from .api import (EveMaterial, DISPLAY_ALBEDO, DISPLAY_DEFAULT, DISPLAY_DENOISE, DISPLAY_DEPTH, DISPLAY_NORMAL, float_buffer_rgb8, write_png,
                  ABSORB, BLEND_ADDITION, BLEND_DARKEN, BLEND_LIGHTEN, BLEND_SUBTRACTION, WRAP_CLAMP, WRAP_REPEAT, Camera, CubeMap, Dielectric,
                  DiffuseLight, FastRand, Instance, Lambertian, Metal, Mix, Model, MrtError, NativeScene, ObjFns, ObjLoader, PlyLoader, Renderer,
                  SimpleTexturedBuilder, SkyBackground,
                  SkySphere, SolidBackground, SolidColor, SolidColorFallback, Specular, Sphere, StlLoader, Texture, TextureBlend, TextureFile, Triangles, V3, V3_fill,
                  Volume, World, YCbCrTexture, render)
from . import scenes  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
