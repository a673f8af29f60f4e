def make_log_joint_fn(model):
    raise NotImplementedError("use the reference's own program_transformations.make_log_joint_fn")
