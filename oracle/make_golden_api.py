"""Test infrastructure (oracle/): records the signature and return structure of the reference's DAgger update,
``HierarchicalTrainer._update_agent`` (robo_vln_baselines/hierarchical_trainer.py:492-560), read with ``ast`` from the
UNMODIFIED reference source, into tests/golden/update_agent_api.json.  Run here (needs /root/reference); the fixture
travels, the reference does not.
    python oracle/make_golden_api.py
"""
import ast
import json
import os

REF = "/root/reference/robo_vln_baselines/hierarchical_trainer.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "update_agent_api.json")

tree = ast.parse(open(REF).read())
fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "_update_agent")
args = [a.arg for a in fn.args.args if a.arg != "self"]
ret = next(n for n in ast.walk(fn) if isinstance(n, ast.Return))
ret_names = [ast.unparse(e) for e in ret.value.elts]
loss = next(n for n in ast.walk(fn) if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "loss" for t in n.targets))
json.dump({"source": "robo_vln_baselines/hierarchical_trainer.py", "function": "_update_agent", "lineno": fn.lineno,
           "args": args, "returns": ret_names, "loss_tuple_len": len(loss.value.elts)}, open(OUT, "w"), indent=1)
print(open(OUT).read())
