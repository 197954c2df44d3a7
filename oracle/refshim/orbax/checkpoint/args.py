class StandardSave:
    def __init__(self, item):
        self.item = item


class StandardRestore:
    def __init__(self, item):
        self.item = item
