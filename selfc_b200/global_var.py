"""Process-global clip length, mirroring the reference's `global_var.GlobalVar` (codes/global_var.py:3-17):
every reshape [B*T,C,H,W] <-> [B,C,T,H,W] on the path reads it, and the dataset constructor sets it
(data/LQGTVID_dataset.py:50)."""


class GlobalVar:
    VIDEO_T_LEN = None
    Istrain = None

    @staticmethod
    def get_Temporal_LEN():
        return GlobalVar.VIDEO_T_LEN

    @staticmethod
    def set_Temporal_LEN(v):
        GlobalVar.VIDEO_T_LEN = v

    @staticmethod
    def get_Istrain():
        return GlobalVar.Istrain

    @staticmethod
    def set_Istrain(v):
        GlobalVar.Istrain = v
