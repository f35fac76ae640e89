"""Build the fitting ``MjModel`` with the real MuJoCo compiler when it is installed.

Mirrors reference ``stac_mjx/stac.py:185-224`` (``MjSpec.from_file`` ->
``add_site`` per keypoint -> ``dm_scale_spec`` -> ``compile``).  Only imported
when ``mujoco`` is importable; this image has no MuJoCo, so the MJCF reader in
`mjcf.py` is what runs here.
"""


def build_mjmodel(mujoco, xml_path, model_cfg):
    spec = mujoco.MjSpec.from_file(str(xml_path))
    size = float(model_cfg.get("MARKER_SIZE", 0.005))
    for key, body_name in model_cfg["KEYPOINT_MODEL_PAIRS"].items():
        pos = model_cfg["KEYPOINT_INITIAL_OFFSETS"][key]
        if isinstance(pos, str):
            pos = [float(p) for p in pos.split(" ")]
        spec.body(body_name).add_site(name=key, size=[size] * 3, rgba=(0, 0, 0, 0.8), pos=pos, group=3)
    scale = float(model_cfg["SCALE_FACTOR"])

    def scale_bodies(parent):
        body = parent.first_body()
        while body:
            body.pos = body.pos * scale
            for geom in body.geoms:
                geom.fromto = geom.fromto * scale
                geom.size = geom.size * scale
                geom.pos = geom.pos * scale
            scale_bodies(body)
            body = parent.next_body(body)

    for mesh in spec.meshes:
        mesh.scale = mesh.scale * scale
    scale_bodies(spec.worldbody.first_body())
    return spec.compile()
