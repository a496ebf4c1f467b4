"""One experiment's recordings on disk (host-side mirror of CEIT/datahandler/Dataset.py, without the plots).

Folder layout and file names are the reference's (Dataset.py:21-46):
    ./data/Smp<sample>_<exp>/<loop>/calibration_sample<..>_ex<..>_<height idx>.csv
    ./data/Smp<sample>_<exp>/<loop>/data_sample<..>_ex<..>_<x>_<y>_<height>.csv
"""
import os

from .Data import rev_translate_corr
from .DataHandler import DataHandler


class DataSet:
    def __init__(self, smp, exp, loop_num, height_list, initialize=True, solver=None):
        self.sample_num, self.exp, self.loop_num = smp, exp, loop_num
        self.height_list = height_list
        self.height_to_idx = {h: i for i, h in enumerate(height_list)}
        self.folder_path = "./data/Smp" + self.sample_num + "_" + self.exp + "/" + str(self.loop_num)
        self.dataset = {h: [] for h in self.height_list}
        self.calibration_data = {}
        self.handler = DataHandler(self.folder_path, solver=solver)
        if initialize:
            self.read_all_files_in_folder()

    def read_all_files_in_folder(self):
        names = []
        for _, _, files in os.walk(self.folder_path):
            names = files
        for name in names:
            if name.split('_')[0] == "calibration":
                one = self.handler.read_one_calibration_file(name, contain_excitation=True)
                if one.height_idx > len(self.height_list) - 1:
                    raise Exception("The height list provided doesn't fit the data. Aborting...")
                self.calibration_data[self.height_list[one.height_idx]] = one.return_copy()
            else:
                one = self.handler.read_one_data_file(name, contain_excitation=True)
                if one.z not in self.height_to_idx.keys():
                    raise Exception("The height list provided doesn't fit the data. Aborting...")
                self.dataset[one.z].append(one.return_copy())

    def __getitem__(self, height):
        if len(self.dataset.keys()) == 0:
            raise Exception(
                "Please call read_all_files_in_folder() function on the object to get data or you have an empty folder")
        if height not in self.height_list:
            raise IndexError("Height not inside height list")
        return {"calibration_data": self.calibration_data[height], "dataset": self.dataset[height]}

    def get_data_at(self, x, y, height):
        for data in self[height]["dataset"]:
            if data.x == x and data.y == y:
                return data.return_copy()
        raise IndexError("x and y not inside dataset")

    def get_calibration_data_at(self, height):
        return self[height]["calibration_data"].return_copy()

    def return_path_for_specific_point(self, x, y, h):
        h_idx = self.height_to_idx[h]
        x, y = rev_translate_corr(x, y)
        stem = "_sample" + self.sample_num + "_ex" + self.exp + "0" + str(self.loop_num) + "_"
        return (self.folder_path, "calibration" + stem + str(h_idx) + ".csv",
                "data" + stem + str(x) + "_" + str(y) + "_" + str(self.height_list[h_idx]) + ".csv")

    def cop_tables(self):
        """The per-height numbers of draw_full_COP_map_for_every_height (Dataset.py:150-166), one batched
        reconstruction per height."""
        return {h: self.handler.cop_table(self.dataset[h], self.calibration_data[h]) for h in self.height_list}

    def draw_full_COP_map_for_every_height(self):
        raise NotImplementedError("plotting is outside the scope of ceit_b200 (the numbers: cop_tables)")
