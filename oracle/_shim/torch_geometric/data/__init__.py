class Data:  # import-only placeholder
    pass


class Batch:  # import-only placeholder
    pass
