/*
 * Drop-in for the three Tools methods on the per-view acquisition path (S = src/main/java/net/preibisch/simulation):
 *
 *     poissonProcess  S/Tools.java:73    in place, lambda = v * (SNR / sqrt 5)^2, raw counts
 *     normImage       S/Tools.java:112   in place, sum becomes 1 (accurate double sum, like mpicbg RealSum)
 *     adjustImage     S/Tools.java:143   in place, two roundings f32( f32( v * corr ) + minValue ), returns corr
 *
 * These are element-wise plus one reduction, so the image travels as a flat list in ITS OWN iteration order and is written
 * back in the same order -- any Iterable< FloatType > works (ArrayImg: bulk copy of the float[]).
 */
package net.preibisch.simulation.gpu;

import java.nio.FloatBuffer;
import java.util.Iterator;
import java.util.Random;

import net.imglib2.IterableInterval;
import net.imglib2.RandomAccessibleInterval;
import net.imglib2.type.numeric.real.FloatType;
import net.imglib2.view.Views;

public class ToolsGPU
{
	private static long count( final Iterable< FloatType > img )
	{
		if ( img instanceof IterableInterval )
			return ( ( IterableInterval< ? > ) img ).size();
		long n = 0;
		for ( final Iterator< FloatType > it = img.iterator(); it.hasNext(); it.next() )
			++n;
		return n;
	}

	private static FloatBuffer toPinned( final Iterable< FloatType > img, final long n )
	{
		final FloatBuffer buf = Mvsim.pinnedFloats( n );
		final float[] array = SimulateMultiViewDatasetGPU.backingArray( img );
		if ( array != null && array.length == n )
			buf.put( array );
		else
			for ( final FloatType t : img )
				buf.put( t.get() );
		buf.rewind();
		return buf;
	}

	private static void copyBack( final FloatBuffer buf, final Iterable< FloatType > img, final long n )
	{
		buf.rewind();
		final float[] array = SimulateMultiViewDatasetGPU.backingArray( img );
		if ( array != null && array.length == n )
			buf.get( array );
		else
			for ( final FloatType t : img )
				t.set( buf.get() );
	}

	/** drop-in for S/Tools.java:73-86 (in place); keyed by one rnd.nextLong() */
	public static void poissonProcess( final RandomAccessibleInterval< FloatType > img, final double SNR, final Random rnd )
	{
		final Iterable< FloatType > flat = Views.flatIterable( img );
		final long n = count( flat );
		final FloatBuffer buf = toPinned( flat, n );
		try
		{
			SimulateMultiViewDatasetGPU.check( Mvsim.poisson( SimulateMultiViewDatasetGPU.ctx(), buf, n, SNR, rnd.nextLong(), 0 ) );
			copyBack( buf, flat, n );
		}
		finally
		{
			Mvsim.freePinned( buf );
		}
	}

	/** drop-in for S/Tools.java:112-118 (in place) */
	final public static void normImage( final Iterable< FloatType > img )
	{
		final long n = count( img );
		final FloatBuffer buf = toPinned( img, n );
		try
		{
			SimulateMultiViewDatasetGPU.check( Mvsim.psfNormalize( SimulateMultiViewDatasetGPU.ctx(), buf, new long[] { n, 1, 1 }, new double[ 1 ] ) );
			copyBack( buf, img, n );
		}
		finally
		{
			Mvsim.freePinned( buf );
		}
	}

	/** drop-in for S/Tools.java:143-159 (in place); returns the factor all intensities were multiplied with */
	public static double adjustImage( final IterableInterval< FloatType > image, final float minValue, final float targetAverage )
	{
		final long n = image.size();
		final FloatBuffer buf = toPinned( image, n );
		final double[] correction = new double[ 1 ];
		try
		{
			SimulateMultiViewDatasetGPU.check( Mvsim.adjust( SimulateMultiViewDatasetGPU.ctx(), buf, new long[] { n, 1, 1 }, minValue, targetAverage, correction ) );
			copyBack( buf, image, n );
		}
		finally
		{
			Mvsim.freePinned( buf );
		}
		return correction[ 0 ];
	}
}
