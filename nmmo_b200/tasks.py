"""Task table: a closed, device-evaluable vocabulary of the predicates the reference uses.

The reference assigns every agent one task drawn from a curriculum of ``TaskSpec`` objects
(/root/reference/reinforcement_learning/environment.py:48-49; ``nmmo.task.task_spec``), whose
``eval_fn`` are Python callables from ``nmmo.task.base_predicates``.  Arbitrary callables cannot
run on the device, so a task here is a row ``int32[12] = {pred, p0, p1, p2, p3, pred2, q0, combine, q1, q2, wa, wb}``
over the predicate names the reference imports
(curriculum_generation/manual_curriculum.py:8-29, neurips23_evaluation/heldout_evaluation_task.py:7-20,
syllabus_wrapper.py:58-70).  ``combine = 1`` is the product form ``a * b`` (manual_curriculum.py:201-202,
``InventorySpaceGE * TickGE``), ``combine = 2`` the weighted sum ``wa/1000 * a + wb/1000 * b``
(manual_curriculum.py:119-122, ``0.3 * EquipItem + 0.7 * GainExperience``); ``b`` takes ``(q0, q1, q2)`` and must
be a state predicate.  Each task also carries a ``task_dim`` fp16 embedding (the
reference's come from an LLM, curriculum_generation/task_encoder.py:87; here they are seeded
synthetic vectors because the curriculum pickle is missing from the reference,
.MISSING_LARGE_BLOBS:1).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .config import SPEC

EVENT = {k[3:]: v for k, v in SPEC.items() if k.startswith("EV_")}
SKILL = {k[3:]: v for k, v in SPEC.items() if k.startswith("SK_")}
ITEM = {k[3:]: v for k, v in SPEC.items() if k.startswith("IT_")}
PRED = {k[3:]: v for k, v in SPEC.items() if k.startswith("TP_")}
NCOL = SPEC["NM_TASK_COLS"]
TF_TEAM, TF_RELATIVE_TARGET = SPEC["TF_TEAM"], SPEC["TF_RELATIVE_TARGET"]


STATE_PREDICATES = ("TICK_GE", "ATTAIN_SKILL", "GAIN_EXPERIENCE", "EQUIP_ITEM", "HOARD_GOLD", "INVENTORY_SPACE_GE", "OWN_ITEM",
                    "FULLY_ARMED", "STAY_ALIVE", "DISTANCE_TRAVELED", "OCCUPY_TILE", "ALL_DEAD", "ALL_MEMBERS_WITHIN_RANGE")


def task_row(pred: str, p0=0, p1=0, p2=0, p3=0, pred2: str = "NONE", q0=0, combine=0, q1=0, q2=0, wa=0, wb=0) -> List[int]:
    """p3 = flags (TF_TEAM | TF_RELATIVE_TARGET, include/nmmo_spec.h nm_task_flag)."""
    if combine and pred2 not in STATE_PREDICATES:
        raise ValueError(f"the second predicate of a combined task must be a state predicate, not {pred2}")
    return [PRED[pred], int(p0), int(p1), int(p2), int(p3), PRED[pred2], int(q0), int(combine), int(q1), int(q2), int(wa), int(wb)]


def practice_skill_with_tool(skill: str, exp: int) -> List[int]:
    """manual_curriculum.py:112-122: 0.3 * EquipItem(tool of the skill, level 1) + 0.7 * GainExperience(skill, exp)."""
    tool = {"MELEE": "SPEAR", "RANGE": "BOW", "MAGE": "WAND", "FISHING": "ROD", "HERBALISM": "GLOVES", "PROSPECTING": "PICKAXE",
            "CARVING": "AXE", "ALCHEMY": "CHISEL"}[skill]      # TOOL_FOR_SKILL, manual_curriculum.py:40-51
    return task_row("EQUIP_ITEM", ITEM[tool], 1, pred2="GAIN_EXPERIENCE", q0=SKILL[skill], q1=exp, combine=2, wa=300, wb=700)


def create_basic_tasks(unit_count: int) -> List[List[int]]:
    """syllabus_wrapper.py:58-70 ``create_basic_tasks``."""
    u = unit_count
    return [
        task_row("TICK_GE", 50 * u),
        task_row("COUNT_EVENT", EVENT["EAT_FOOD"], 5 * u),
        task_row("COUNT_EVENT", EVENT["DRINK_WATER"], 5 * u),
        task_row("COUNT_EVENT", EVENT["HARVEST_ITEM"], 3 * u),
        task_row("COUNT_EVENT", EVENT["GO_FARTHEST"], 3 * u),
        task_row("COUNT_EVENT", EVENT["LEVEL_UP"], 2 * u),
        task_row("COUNT_EVENT", EVENT["EQUIP_ITEM"], u),
        task_row("COUNT_EVENT", EVENT["CONSUME_ITEM"], u),
        task_row("COUNT_EVENT", EVENT["BUY_ITEM"], u),
        task_row("COUNT_EVENT", EVENT["PLAYER_KILL"], u),
    ]


def default_curriculum() -> List[List[int]]:
    """A curriculum exercising every predicate family of manual_curriculum.py:53-314."""
    rows: List[List[int]] = []
    for u in (1, 2, 3, 4, 5):                      # syllabus_wrapper.py:36-55 five stages
        rows += create_basic_tasks(u)
    for mat in (SPEC["MT_FOILAGE"], SPEC["MT_WATER"], SPEC["MT_ORE"], SPEC["MT_TREE"], SPEC["MT_HERB"]):
        rows.append(task_row("CAN_SEE_TILE", mat))
    for skill in SKILL.values():
        rows.append(task_row("ATTAIN_SKILL", skill, 3))
        rows.append(task_row("GAIN_EXPERIENCE", skill, 50))
    for style in (1, 2, 3):
        rows.append(task_row("SCORE_HIT", style, 10))
        rows.append(task_row("FULLY_ARMED", style, 1))
    for amount in (5, 20):
        rows.append(task_row("HOARD_GOLD", amount))
        rows.append(task_row("EARN_GOLD", amount))
        rows.append(task_row("SPEND_GOLD", amount))
        rows.append(task_row("MAKE_PROFIT", amount))
    for it in ("RATION", "POTION", "WHETSTONE", "ARROW", "RUNES", "SPEAR"):
        rows.append(task_row("HARVEST_ITEM", ITEM[it], 1, 3))
        rows.append(task_row("OWN_ITEM", ITEM[it], 1, 2))
        rows.append(task_row("LIST_ITEM", ITEM[it], 1, 1))
        rows.append(task_row("BUY_ITEM", ITEM[it], 1, 1))
    rows.append(task_row("CONSUME_ITEM", ITEM["RATION"], 1, 2))
    rows.append(task_row("CONSUME_ITEM", ITEM["POTION"], 1, 2))
    for it in ("HAT", "TOP", "BOTTOM", "ROD", "PICKAXE"):
        rows.append(task_row("EQUIP_ITEM", ITEM[it], 1))
    rows.append(task_row("DEFEAT_ENTITY", 0, 1, 2))
    rows.append(task_row("DEFEAT_ENTITY", 1, 1, 1))
    rows.append(task_row("INVENTORY_SPACE_GE", 6, pred2="TICK_GE", q0=200, combine=1))
    rows.append(task_row("DISTANCE_TRAVELED", 32))
    rows.append(task_row("STAY_ALIVE"))
    rows.append(task_row("OCCUPY_TILE", 80, 80))
    rows.append(task_row("CAN_SEE_AGENT", 1))
    rows.append(task_row("CAN_SEE_GROUP", 1, 8))
    rows.append(task_row("ALL_MEMBERS_WITHIN_RANGE", 5, pred2="TICK_GE", q0=100, combine=1))
    for skill in ("MELEE", "FISHING", "PROSPECTING"):
        rows.append(practice_skill_with_tool(skill, 30))
    return rows


def make_task_table(rows: Sequence[Sequence[int]], task_dim: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """-> (int32[T, NM_TASK_COLS = 12] specs, uint16[T, task_dim] fp16 bit patterns of the embeddings)."""
    tab = np.asarray(rows, np.int32).reshape(-1, NCOL)
    rng = np.random.default_rng(seed)
    emb = (rng.standard_normal((tab.shape[0], task_dim)) * 0.5).astype(np.float16)
    return np.ascontiguousarray(tab), np.ascontiguousarray(emb.view(np.uint16))
