"""The render configurations of BASELINE.json / SURVEY.md §8d."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class RenderConfig:
    name: str
    width: int
    height: int
    aa: int
    shadow_samples: int
    max_bounces: int
    description: str

    @property
    def focal(self) -> float:
        # the reference couples f = 2200 to aa = 2, H = 1024 (skeleton.cpp:61): keep the box fitted
        return 1100.0 * self.aa * self.height / 1024.0


CONFIGS = {
    "head": RenderConfig("head", 1024, 1024, 2, 10, 10, "reference HEAD defaults: 1024^2, 2x2 AA, 10 shadow samples, <=10 bounces"),
    "cfg1": RenderConfig("cfg1", 1024, 1024, 1, 1, 0, "Cornell Box, default resolution, 1 spp, direct light only"),
    "cfg2": RenderConfig("cfg2", 1920, 1080, 2, 8, 10, "Cornell Box 1080p, reflection/refraction, 8-sample soft shadows, 4x AA"),
    "cfg3": RenderConfig("cfg3", 3840, 2160, 4, 10, 4, "Cornell Box 4K, 16 spp AA, soft shadows, 4 bounces"),
    "cfg4": RenderConfig("cfg4", 1920, 1080, 2, 8, 10, "Loader.cpp synthetic 1.31 M-triangle OBJ mesh (icosphere, 8 subdivisions) inside the Cornell Box, GPU-built BVH, 1080p"),
    "cfg5": RenderConfig("cfg5", 7680, 4320, 2, 10, 10, "8K frame of the full-feature Cornell Box"),
}
