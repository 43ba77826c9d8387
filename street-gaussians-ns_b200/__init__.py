"""street-gaussians-ns_b200: B200-native (sm_100a) differentiable Gaussian rasterizer hot path.

Drop-in for the gsplat-0.1.x calls made by street_gaussians_ns/sgn_splatfacto.py:860-873,939,954-994
and the scene-graph compose of street_gaussians_ns/sgn_splatfacto_scene_graph.py:305-374.
See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1.0"
