class ModelSummary:
    def __init__(self, *a, **k):
        pass

    def __str__(self):
        return "ModelSummary(refshim)"
