"""API constants re-stated from the reference (values are part of the wire format).

ConeTypes:    fsd_path_planning/utils/cone_types.py:10-19
MissionTypes: fsd_path_planning/utils/mission_types.py:11-25
"""
from enum import IntEnum


class ConeTypes(IntEnum):
    UNKNOWN = 0
    RIGHT = YELLOW = 1
    LEFT = BLUE = 2
    START_FINISH_AREA = ORANGE_SMALL = 3
    START_FINISH_LINE = ORANGE_BIG = 4


class MissionTypes(IntEnum):
    none = 0
    acceleration = 1
    skidpad = 2
    autocross = 3
    trackdrive = 4
    ebs_test = 5
    inspection = 6
    manual_driving = 7
