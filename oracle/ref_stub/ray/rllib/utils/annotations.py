def override(cls):
    def check_override(method):
        return method
    return check_override
