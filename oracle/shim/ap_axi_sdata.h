// ap_axi_sdata.h -- stand-in for the Xilinx AXI4-Stream side-channel struct.
// TEST INFRASTRUCTURE ONLY (see ap_int.h).  ap_axiu<D,U,TI,TD>: data D bits,
// keep/strb D/8 bits, last 1 bit; zero-width side channels are modelled as 1 bit.
#ifndef FR_SHIM_AP_AXI_SDATA_H
#define FR_SHIM_AP_AXI_SDATA_H
#include "ap_int.h"
template <int D, int U, int TI, int TD>
struct ap_axiu {
  ap_uint<D> data;
  ap_uint<(D + 7) / 8> keep;
  ap_uint<(D + 7) / 8> strb;
  ap_uint<(U > 0 ? U : 1)> user;
  ap_uint<1> last;
  ap_uint<(TI > 0 ? TI : 1)> id;
  ap_uint<(TD > 0 ? TD : 1)> dest;
};
#endif
