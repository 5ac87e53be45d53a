"""Stand-in for py_expression_eval.Parser (test infrastructure, see README.md): upstream only calls
``Parser().parse(s).variables()`` and ``.evaluate(dict)`` on conditions such as ``"kernel == 'rbf'"``."""
import ast


class _Expression:
    def __init__(self, text):
        self.text = text
        self.tree = ast.parse(text.strip(), mode="eval")

    def variables(self):
        seen = []
        for node in ast.walk(self.tree):
            if isinstance(node, ast.Name) and node.id not in seen and node.id not in ("True", "False", "None"):
                seen.append(node.id)
        return seen

    def evaluate(self, values):
        return eval(compile(self.tree, "<condition>", "eval"), {"__builtins__": {}}, dict(values))


class Parser:
    def parse(self, text):
        return _Expression(text)
