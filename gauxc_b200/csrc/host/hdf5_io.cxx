#include "hdf5_io.hpp"

namespace GauXC {

void read_hdf5_record(Molecule&, const std::string&, const std::string&) {
  GAUXC_GENERIC_EXCEPTION("HDF5 Molecule read NYI in B200 path");
}
void read_hdf5_record(BasisSet&, const std::string&, const std::string&) {
  GAUXC_GENERIC_EXCEPTION("HDF5 BasisSet read NYI in B200 path");
}
void write_hdf5_record(const Molecule&, const std::string&, const std::string&) {
  GAUXC_GENERIC_EXCEPTION("HDF5 Molecule write NYI in B200 path");
}
void write_hdf5_record(const BasisSet&, const std::string&, const std::string&) {
  GAUXC_GENERIC_EXCEPTION("HDF5 BasisSet write NYI in B200 path");
}
void read_hdf5_dataset(const std::string&, const std::string&, std::vector<double>&, std::vector<size_t>&) {
  GAUXC_GENERIC_EXCEPTION("HDF5 dataset read NYI in B200 path");
}
void write_hdf5_dataset(const std::string&, const std::string&, const double*, const std::vector<size_t>&) {
  GAUXC_GENERIC_EXCEPTION("HDF5 dataset write NYI in B200 path");
}

}  // namespace GauXC
