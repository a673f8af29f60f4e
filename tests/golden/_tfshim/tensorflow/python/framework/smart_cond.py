smart_cond = None
